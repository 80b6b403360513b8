#!/bin/bash
# One gpurun call = several measurements; every step logs under gpurun_out/ and no step aborts the others.
# usage: tools/gpu_session.sh <tag> <step> [<step> ...]     steps: tests bench bench_c3 bench_c5 parity profile
tag=$1; shift
mkdir -p gpurun_out
for step in "$@"; do
  case $step in
    tests)    timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log ;;
    bench)    timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.log 2>&1; echo "exit $?" >> gpurun_out/${tag}_bench.log ;;
    benchq)   timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-strong-c4 --no-e2e --no-models > gpurun_out/${tag}_benchq.log 2>&1; echo "exit $?" >> gpurun_out/${tag}_benchq.log ;;
    bench_c3) timeout 600 python bench.py --workload c3 --steps 5 --warmup 3 > gpurun_out/${tag}_bench_c3.log 2>&1; echo "exit $?" >> gpurun_out/${tag}_bench_c3.log ;;
    bench_c5) timeout 600 python bench.py --workload c5 --steps 4 --warmup 3 > gpurun_out/${tag}_bench_c5.log 2>&1; echo "exit $?" >> gpurun_out/${tag}_bench_c5.log ;;
    parity)   timeout 1800 python tools/parity_full.py c2 c3 c5 c4n1 > gpurun_out/${tag}_parity.log 2>&1; echo "exit $?" >> gpurun_out/${tag}_parity.log ;;
    profile)  timeout 600 python tools/profile_fit.py c2 > gpurun_out/${tag}_profile.log 2>&1; echo "exit $?" >> gpurun_out/${tag}_profile.log ;;
    kern)     { timeout 300 python tools/bench_kernels.py 8760 1038240 60 --env=XEOFS_TC_CLUSTER=1,2,4,1,2;
                timeout 300 python tools/bench_kernels.py 8760 518400 110 --env=XEOFS_TC_CLUSTER=1,2,4;
                timeout 300 python tools/bench_kernels.py 8760 259200 30 --env=XEOFS_TC_CLUSTER=1,2,4; } > gpurun_out/${tag}_kern.log 2>&1 ;;
    clchk)    XEOFS_TC_CLUSTER=2 timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "tcgen05 or fused" > gpurun_out/${tag}_clchk.log 2>&1; echo "exit $?" >> gpurun_out/${tag}_clchk.log ;;
    small)    { timeout 300 python tools/bench_small.py; timeout 300 python tools/bench_small.py 518400 110; } > gpurun_out/${tag}_small.log 2>&1 ;;
    profile_c4) timeout 900 python tools/profile_fit.py c4 > gpurun_out/${tag}_profile_c4.log 2>&1 ;;
    vt)       { timeout 150 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "varimax or gram"; echo "pytest exit $?"; timeout 150 python -m pytest tests/test_gpu_models.py -x -q -m gpu -k "rotat"; echo "pytest exit $?";
                timeout 120 python tools/bench_varimax.py; } > gpurun_out/${tag}_vt.log 2>&1 ;;
    eig)      timeout 600 python tools/diag_eig.py > gpurun_out/${tag}_eig.log 2>&1 ;;
    c5tol)    { for v in 1e-9 1e-8 3e-8; do echo "== XEOFS_TC_EIG_TOL=$v"; XEOFS_TC_EIG_TOL=$v timeout 300 python bench.py --workload c5 --steps 4 --warmup 3 --no-cpu; done; } > gpurun_out/${tag}_c5tol.log 2>&1 ;;
    profile_c3) timeout 600 python tools/profile_fit.py c3 > gpurun_out/${tag}_profile_c3.log 2>&1 ;;
    profile_c5) timeout 600 python tools/profile_fit.py c5 > gpurun_out/${tag}_profile_c5.log 2>&1 ;;
    two)      { timeout 600 python -m pytest tests -m gpu -q -k "two_gpus"; echo "pytest exit $?";
                timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3; echo "exit $?"; } > gpurun_out/${tag}_two.log 2>&1 ;;
    eight)    { timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3; echo "exit $?"; } > gpurun_out/${tag}_eight.log 2>&1 ;;
    four)     { timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 10 --warmup 3; echo "exit $?"; } > gpurun_out/${tag}_four.log 2>&1 ;;
    refs8)    timeout 900 python bench.py --impl reference --ref-cols 129780 --steps 1 --warmup 0 > gpurun_out/${tag}_refs8.log 2>&1 ;;
    c4n1)     timeout 900 python tools/parity_full.py c4n1 > gpurun_out/${tag}_c4n1.log 2>&1 ;;
    smoke)    timeout 300 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "exit $?" >> gpurun_out/${tag}_smoke.log ;;
    *) echo "unknown step $step" ;;
  esac
done
tail -n 3 gpurun_out/${tag}_*.log | cut -c1-600
