for v in 0 1 2 3; do echo "K10 variant $v"; HMC_B200_K10_VARIANT=$v python scripts/probe_c2.py 2000 1000 0.05 10 128 2>&1 | tail -1; done
HMC_B200_K10_VARIANT=0 python scripts/probe_c2.py 2000 1000 0.05 10 64,256 2>&1 | tail -2
