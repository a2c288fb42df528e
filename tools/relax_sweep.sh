for w in 100000; do for t in 48; do echo "wide=$w tile=$t"; MLP_RELAX_WIDE=$w MLP_RELAX_TILE=$t timeout 120 python tools/prof_run.py 192 300 qp 2>&1 | grep "relax wall" | tail -1; done; done
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "qp_against or ragged or cpnp_against" 2>&1 | tail -3
