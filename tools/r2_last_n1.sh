#!/bin/bash
# last single-GPU evidence of the round (lean: the reference arm, CLI walls and scene-creation profile of r2z stand)
T=${1:-r2f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${T}_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/${T}_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/${T}_launches_bench.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 2 -c 1 -o gpurun_out/${T}_render_1920 -f python tools/ncu_small_frame.py horse_and_mug 16 1920 960 > gpurun_out/${T}_ncu_1920.log 2>&1; echo "ncu1 rc=$?"
ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section SchedulerStats --section WarpStateStats --section InstructionStats --section Occupancy --section LaunchStats --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_pred_on_per_inst_executed.ratio,smsp__thread_inst_executed_per_inst_executed.ratio,sm__issue_active.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_write.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum --clock-control none -k regex:render_kernel -s 1 -c 1 -o gpurun_out/${T}_render_8k -f python tools/ncu_small_frame.py horse_and_mug 16 7680 3840 > gpurun_out/${T}_ncu_8k.log 2>&1; echo "ncu2 rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','render_kernel_ms')}, d['e2e']['ms_per_frame'], d['scene_build']['cold_s'], d['scene_build']['warm_s'], d['clocks'], d['frame']['frame_sha256'][:16])
PY
