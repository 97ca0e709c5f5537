for v in 4 5 6; do echo "K4 MINB $v"; HMC_B200_K4_MINB=$v python scripts/probe_c2.py 2000 1000 0.05 4 128 2>&1 | tail -1; done
