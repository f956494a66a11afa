#!/bin/bash
# round 2, call B: new parity tests (with their printed measurements), CTA size x lanes A/B, NODE_STEPS=2 variant, Cornell lanes,
# bench line, ncu launch list + full captures (raw CSV made on the box: .ncu-rep files over 64 MiB do not come back)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider -k "shade_step or crop or text or lanes or async or host_app or roulette or next_event or front_end or sah or build_invariants or soup_build or ring" > $O/r2b_pytest.txt 2>&1
echo "pytest exit $?" >> $O/r2b_pytest.txt
{
for blk in 1024 256; do for lanes in 1 2; do
  timeout 300 python tools/probe.py --size 4096 --spp 8 --frames 3 --opt 4=$lanes --opt 5=$blk --tag full_b${blk}_l$lanes
  timeout 300 python tools/probe.py --size 4096 --spp 8 --frames 3 --tile 8,0 --opt 4=$lanes --opt 5=$blk --tag tile8_b${blk}_l$lanes
done; done
BPT_LIB_VARIANT=ns2 timeout 300 python tools/probe.py --size 4096 --spp 8 --frames 3 --opt 4=1 --tag ns2_full
BPT_LIB_VARIANT=ns2 timeout 300 python tools/probe.py --size 4096 --spp 8 --frames 3 --tile 8,0 --opt 4=1 --tag ns2_tile8
timeout 300 python tools/probe.py --size 4096 --spp 8 --frames 3 --opt 4=1 --opt 13=0 --tag noring_full
timeout 300 python tools/probe.py --size 4096 --spp 8 --frames 3 --opt 4=1 --opt 12=0 --tag sah0_full
timeout 300 python tools/probe.py --size 4096 --spp 8 --frames 3 --opt 4=1 --opt 12=8 --tag sah8_full
timeout 300 python tools/probe.py --tris 1000000 --seed 0x5EED0001 --size 2048 --spp 8 --frames 3 --opt 4=1 --opt 12=0 --tag sah0_soup1m
timeout 300 python tools/probe.py --tris 1000000 --seed 0x5EED0001 --size 2048 --spp 8 --frames 3 --opt 4=1 --tag sah32_soup1m
} > $O/r2b_probe.txt 2>&1
for lanes in 1 2; do
  timeout 300 python bench.py --workload cornell --no-e2e --no-cpu-baseline --opt 4=$lanes > $O/r2b_bench_cornell_l$lanes.json 2>> $O/r2b_bench.err
  timeout 300 python bench.py --workload cornell1000 --no-e2e --no-cpu-baseline --opt 4=$lanes > $O/r2b_bench_cornell1000_l$lanes.json 2>> $O/r2b_bench.err
done
timeout 300 python bench.py --workload cornell --no-e2e --no-cpu-baseline --opt 4=1 --opt 13=0 > $O/r2b_bench_cornell_noring.json 2>> $O/r2b_bench.err
timeout 600 python bench.py --steps 8 --warmup 3 --opt 4=1 > $O/r2b_bench_soup10m.json 2>> $O/r2b_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2b_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --opt 4=1 > $O/r2b_bench_under_ncu.log 2>&1
# the 8 traversal launches (one lane, bounces 0..7) of the timed frame; launches 0..7 are the warm-up frame
timeout 900 ncu --set full --clock-control none -k regex:k_trace -s 8 -c 8 -f -o /tmp/r2b_k_trace_all \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --opt 4=1 > $O/r2b_ncu_trace.log 2>&1
ncu -i /tmp/r2b_k_trace_all.ncu-rep --page raw --csv > $O/r2b_k_trace_all_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 9 -c 2 -f -o $O/r2b_k_trace_src \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --opt 4=1 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 9 -c 2 -f -o $O/r2b_k_shade \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --opt 4=1 > $O/r2b_ncu_shade.log 2>&1
du -sh $O; ls -la $O
tail -30 $O/r2b_pytest.txt
cat $O/r2b_probe.txt
