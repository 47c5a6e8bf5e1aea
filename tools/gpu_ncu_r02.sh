#!/bin/bash
# Round-2 ncu evidence (run on a B200: gpurun -- 'bash tools/gpu_ncu_r02.sh'); condensed by tools/summarize_profiles.py r02.
# gpurun copies back at most 64 MiB: every .ncu-rep is exported to a raw-metrics CSV and deleted on the box.
mkdir -p gpurun_out
NCU="ncu --clock-control none"
raw() { ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null; rm -f gpurun_out/$1.ncu-rep; wc -l gpurun_out/$1_raw.csv; }
echo "=== (1) launch list of the default bench command (c3, CUDA graphs, 2 lanes)"
timeout 900 $NCU --metrics gpu__time_duration.sum -c 7000 --csv --log-file gpurun_out/launches_r02_default.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_list_default.log 2>&1; tail -1 gpurun_out/ncu_list_default.log | cut -c1-160
B="python bench.py --dtype bf16 --workload c3 --utts 64 --steps 1 --warmup 1 --lanes 1 --no-graph --no-cpu-baseline --no-extras"
echo "=== (2) --set full captures (c3, 64 utterances, eager launches)"
timeout 600 $NCU --set full -k regex:gemm_tc -s 60 -c 12 -f -o gpurun_out/prof_r02_gemm $B > gpurun_out/ncu_gemm.log 2>&1; raw prof_r02_gemm
timeout 600 $NCU --set full -k regex:attention_tc -s 4 -c 2 -f -o gpurun_out/prof_r02_attn $B > gpurun_out/ncu_attn.log 2>&1; raw prof_r02_attn
timeout 600 $NCU --set full -k regex:conv0_tc -s 1 -c 1 -f -o gpurun_out/prof_r02_conv0 $B > gpurun_out/ncu_conv0.log 2>&1; raw prof_r02_conv0
timeout 600 $NCU --set full -k regex:layernorm -s 6 -c 4 -f -o gpurun_out/prof_r02_ln $B > gpurun_out/ncu_ln.log 2>&1; raw prof_r02_ln
timeout 600 $NCU --set full -k regex:posconv_stacked -s 1 -c 1 -f -o gpurun_out/prof_r02_posconv $B > gpurun_out/ncu_posconv.log 2>&1; raw prof_r02_posconv
echo "=== (3) DRAM traffic per launch (c3, 24 utterances)"
timeout 900 $NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -s 700 -c 1400 --csv --log-file gpurun_out/traffic_r02_c3.csv \
  python bench.py --dtype bf16 --workload c3 --utts 24 --steps 1 --warmup 1 --lanes 1 --no-graph --no-cpu-baseline --no-extras > gpurun_out/ncu_traffic.log 2>&1; tail -1 gpurun_out/ncu_traffic.log | cut -c1-100
echo "=== (4) c5 training step: launch list + captures of the new kernels"
timeout 900 $NCU --metrics gpu__time_duration.sum -c 5000 --csv --log-file gpurun_out/launches_r02_c5.csv \
  python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list_c5.log 2>&1; tail -1 gpurun_out/ncu_list_c5.log | cut -c1-160
timeout 600 $NCU --set full -k regex:"attn_bwd_softmax|transpose64|head_pack|layernorm_bwd|colsum|act_bwd" -s 300 -c 12 -f -o gpurun_out/prof_r02_c5 \
  python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c5.log 2>&1; raw prof_r02_c5
du -sh gpurun_out; ls -la gpurun_out | head -30
