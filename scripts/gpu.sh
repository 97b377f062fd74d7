#!/bin/bash
# One runner for everything that goes to the GPU box:   gpurun -- 'bash scripts/gpu.sh <task> [args]'
# Every task writes under gpurun_out/ (merged back by gpurun) and prints a short tail.
set -u
mkdir -p gpurun_out
export SPZB200_NO_REBUILD=1
task=${1:-tests}; shift || true
case "$task" in
  box)      { nproc; free -g; cat /sys/fs/cgroup/memory.max 2>/dev/null; nvidia-smi -L; nvidia-smi topo -m; lscpu | head -24; numactl -H 2>/dev/null; } > gpurun_out/box.txt 2>&1; cat gpurun_out/box.txt ;;
  smoke)    timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log ;;
  tests)    timeout 2400 python -m pytest tests -m gpu -x -q "$@" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log ;;
  bench)    # bench [name] [bench.py args...]
            name=${1:-bench}; shift || true
            timeout 1500 python bench.py "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "bench rc=$?"; cat gpurun_out/$name.json; tail -3 gpurun_out/$name.err ;;
  benchn)   # benchn N [name] [bench.py args...]: N ranks under torchrun
            n=$1; name=${2:-bench_n$1}; shift 2 || true
            timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@" \
              > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "bench x$n rc=$?"; grep '^{' gpurun_out/$name.json | tail -1; tail -3 gpurun_out/$name.err ;;
  refarm)   timeout 900 python bench.py --impl reference "$@" > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json ;;
  coldstart) timeout 300 python scripts/cold_start.py > gpurun_out/cold_start.txt 2>&1; echo "rc=$?"; cat gpurun_out/cold_start.txt ;;
  link)     # link [args to scripts/link_probe.py]
            timeout 900 python scripts/link_probe.py "$@" > gpurun_out/link_probe.jsonl 2> gpurun_out/link_probe.err; echo "rc=$?"; cat gpurun_out/link_probe.jsonl; tail -3 gpurun_out/link_probe.err ;;
  pool)     timeout 900 python scripts/pool_sweep.py > gpurun_out/pool_sweep.jsonl 2>&1; echo "rc=$?"; cat gpurun_out/pool_sweep.jsonl ;;
  sweep)    # sweep <name> <script> [args]: any scripts/*.py measurement, output kept under gpurun_out/<name>
            name=$1; script=$2; shift 2 || true
            timeout 1500 python scripts/$script "$@" > gpurun_out/$name 2> gpurun_out/$name.err; echo "rc=$?"; cat gpurun_out/$name; tail -3 gpurun_out/$name.err ;;
  abr1)     timeout 1500 bash scripts/ab_vs_r1.sh "$@" > gpurun_out/ab_vs_r1.jsonl 2> gpurun_out/ab_vs_r1.err; echo "rc=$?"; cat gpurun_out/ab_vs_r1.jsonl; tail -3 gpurun_out/ab_vs_r1.err ;;
  api)      # wall time of the C++ drop-in API on pageable std::vector planes (scripts/api_timing.cc); rep 0 includes context creation and pinned buffers
            mkdir -p scripts/_build
            g++ -std=c++17 -O2 -pthread -Iinclude/spz scripts/api_timing.cc -o scripts/_build/api_timing -Lspz_b200/_lib -lspz_b200 -Wl,-rpath,'$ORIGIN/../../spz_b200/_lib' || exit 1
            { for n in 6e4 2e5 1e6 2e6 4e6 1e7; do echo "== $n gaussians SH3 (default policy)"; scripts/_build/api_timing $n 4; done
              echo "== 1e7, SPZ_B200_ZEROFILL=1 (plain resize)"; SPZ_B200_ZEROFILL=1 scripts/_build/api_timing 1e7 4
              echo "== 1e6, SPZ_B200_UNPACK_READAHEAD=0 (every unpack(i, c) is its own launch)"; SPZ_B200_UNPACK_READAHEAD=0 scripts/_build/api_timing 1e6 2
              for n in 1e6 1e7; do echo "== $n, SPZB200_NT_COPY=0 (plain memcpy into / out of the bounce buffers)"; SPZB200_NT_COPY=0 scripts/_build/api_timing $n 4; done; } > gpurun_out/cxx_api_timing.txt 2>&1
            echo "rc=$?"; cat gpurun_out/cxx_api_timing.txt ;;
  smallapi) # small pageable clouds through the C++ API: default staging policy against always-bounce (needs scripts/_build/api_timing: run `api` once, or build here)
            export SPZ_B200_UNPACK_READAHEAD=0
            { for n in 2e4 6e4 1e5; do
                echo "== $n default"; scripts/_build/api_timing $n 6 | grep packGaussians | tail -3
                echo "== $n SPZB200_BOUNCE_MIN_MB=1000 (driver-staged copies)"; SPZB200_BOUNCE_MIN_MB=1000 scripts/_build/api_timing $n 6 | grep packGaussians | tail -3
              done; } > gpurun_out/small_api.txt 2>&1; cat gpurun_out/small_api.txt ;;
  pageable) # copy threads and range size of the bounced pipeline, 10M SH3 through the C++ API
            export SPZ_B200_UNPACK_READAHEAD=0
            run() { echo "== $*"; env "$@" | grep packGaussians | tail -2; }
            { for n in 6e4 2e5 1e6 4e6 1e7; do run scripts/_build/api_timing $n 5; done
              for t in 4 8 12 16; do run SPZB200_COPY_THREADS=$t scripts/_build/api_timing 1e7 4; done
              for c in 131072 524288 1048576; do run SPZB200_PAGEABLE_CHUNK_POINTS=$c scripts/_build/api_timing 1e7 4; done
              run SPZB200_NT_COPY=0 scripts/_build/api_timing 1e7 4; } > gpurun_out/pageable_api.txt 2>&1; cat gpurun_out/pageable_api.txt ;;
  fileapi)  # saveSpz / loadSpz / pack / unpack through the C++ API: this library against the reference's own sources on the same box.
            # scripts/_build/file_api_timing_ref is built where /root/reference exists (see the header of scripts/file_api_timing.cc) and travels with the snapshot.
            g++ -std=c++17 -O2 -pthread -Iinclude/spz scripts/file_api_timing.cc -o scripts/_build/file_api_timing -Lspz_b200/_lib -lspz_b200 -Wl,-rpath,'$ORIGIN/../../spz_b200/_lib' || exit 1
            { for n in 6e4 1e6 4e6; do
                [ -x scripts/_build/file_api_timing_ref ] && scripts/_build/file_api_timing_ref $n 1 reference | grep -v "^\[SPZ"
                scripts/_build/file_api_timing $n 3 spz_b200 | grep -v "^\[SPZ" | tail -2
              done
              scripts/_build/file_api_timing 1e7 2 spz_b200 | grep -v "^\[SPZ" | tail -1
              SPZ_B200_GZIP_THREADS=1 scripts/_build/file_api_timing 1e6 2 "spz_b200 SPZ_B200_GZIP_THREADS=1" | tail -1; } > gpurun_out/file_api.jsonl 2>&1; cat gpurun_out/file_api.jsonl ;;
  hostprobe) # what the host copies per second (threads, memcpy vs non-temporal) and the two shapes of the bounce stage, bare
            nvcc -O2 -Wno-deprecated-gpu-targets -o scripts/_build/host_copy_probe scripts/host_copy_probe.cu -Xcompiler -pthread,-mavx2 || exit 1
            nvcc -O2 -Wno-deprecated-gpu-targets -o scripts/_build/bounce_probe scripts/bounce_probe.cu -Xcompiler -pthread,-mavx2 || exit 1
            timeout 300 scripts/_build/host_copy_probe 1.5e9 > gpurun_out/host_copy_probe.jsonl 2>&1; timeout 400 scripts/_build/bounce_probe 2.36e9 > gpurun_out/bounce_probe.jsonl 2>&1
            cat gpurun_out/host_copy_probe.jsonl gpurun_out/bounce_probe.jsonl ;;
  launches) # launch list of one short bench run (after the same command ran clean without ncu)
            timeout 900 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-ply "$@" > gpurun_out/launch_pre.json 2> gpurun_out/launch_pre.err || { echo "plain run failed"; exit 1; }
            timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'Tiles|PerGaussian|Generic|unpackRecords|Ply|Tables|probePack' -c 400 --csv --log-file gpurun_out/launches.csv \
              python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-ply "$@" > gpurun_out/ncu_list.log 2>&1; echo "ncu rc=$?"; tail -5 gpurun_out/launches.csv ;;
  ncufull)  # ncufull <name> <kernel regex> <skip> <count> <command...>: one ncu --set full capture, after the same command ran clean without ncu
            name=$1; regex=$2; skip=$3; count=$4; shift 4
            timeout 900 "$@" > gpurun_out/${name}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${name}_plain.log; exit 1; }
            timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$regex" -s $skip -c $count -o gpurun_out/$name -f "$@" > gpurun_out/${name}_ncu.log 2>&1
            echo "ncu rc=$?"; grep -c "Profiling" gpurun_out/${name}_ncu.log; ls -la gpurun_out/$name.ncu-rep ;;
  *)        echo "unknown task $task"; exit 2 ;;
esac
