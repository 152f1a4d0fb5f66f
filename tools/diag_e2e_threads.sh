for t in 0 4 8 12 14; do echo threads $t; CTK_WIDEN_THREADS=$t python tools/diag_e2e.py 2>&1 | grep "default"; done
