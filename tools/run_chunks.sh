for ch in 4 8 12 16 32; do echo "chunk $ch"; MMPGO_TS_CHUNK=$ch timeout 60 python tools/tsolve_stress.py 100,100,100 64 3 2>&1 | tail -1; done
