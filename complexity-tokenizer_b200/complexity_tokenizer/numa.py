"""Host placement next to a GPU (no libnuma needed: sysfs + sched_setaffinity).

A rank that feeds one GPU should run on the CPUs of the NUMA node the GPU hangs off: its pages (page-locked input and
result buffers, first touched by this process) then sit on that node and the DMA engines do not cross the socket
interconnect.  The library does this itself for the threads and buffers it owns (csrc/host_api.cu); this module is for
the CALLER's process (bench.py binds each torchrun rank with it)."""
import os

from . import _lib


def node_of_device(device):
    """NUMA node of the host memory next to a CUDA device, -1 if the platform does not say"""
    return int(_lib().ctk_device_numa_node(int(device)))


def cpus_of_node(node):
    if node < 0:
        return set()
    try:
        with open('/sys/devices/system/node/node%d/cpulist' % node) as f:
            spec = f.read().strip()
    except OSError:
        return set()
    cpus = set()
    for part in spec.split(','):
        if not part:
            continue
        a, _, b = part.partition('-')
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def describe(device):
    node = node_of_device(device)
    return {'device': int(device), 'numa_node': node, 'bound': False, 'cpus_allowed': len(os.sched_getaffinity(0))}


def bind_process_to_device(device):
    """run this process on the CPUs next to `device` (within what it is allowed to use); returns what was done"""
    node = node_of_device(device)
    allowed = os.sched_getaffinity(0)
    want = cpus_of_node(node) & allowed
    info = {'device': int(device), 'numa_node': node, 'bound': False, 'cpus_allowed': len(allowed)}
    if want and want != allowed:
        try:
            os.sched_setaffinity(0, want)
            info.update(bound=True, cpus_bound=len(want))
        except OSError as ex:
            info['error'] = repr(ex)
    return info


_INITIAL = None


def unbind_process():
    """back to every CPU the process may use"""
    try:
        os.sched_setaffinity(0, range(os.cpu_count() or 1))
    except OSError:
        pass
