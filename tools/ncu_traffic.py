#!/usr/bin/env python3
"""DRAM bytes of one kernel launch from an `ncu --set full` report -> profiles/traffic.json (read by bench.py).
usage: python tools/ncu_traffic.py report.ncu-rep kernel_name input_bytes 'source note'"""
import csv, io, json, os, subprocess, sys
rep, kernel, nbytes, note = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); h, u, v = rows[0], rows[1], rows[2]
def val(k):
    i = h.index(k); x = float(v[i].replace(',', '')); unit = u[i].lower()
    return x * {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}[unit]
tot = val('dram__bytes_read.sum') + val('dram__bytes_write.sum')
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', 'traffic.json')
d = json.load(open(path)) if os.path.exists(path) else {}
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
d[kernel] = {'dram_bytes_per_launch': int(tot), 'input_bytes': nbytes, 'source': note, 'kernel_source_sha': bench.kernel_source_sha()}
json.dump(d, open(path, 'w'), indent=1)
json.dump(d, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'bench_traffic.json'), 'w'), indent=1)   # copy that travels with the repo snapshot
print(kernel, int(tot), 'bytes per launch')
