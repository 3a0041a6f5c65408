set -x
CMD="python bench.py --steps 2 --warmup 3 --workloads none --no-cpu-baseline"
# 1. every launch of the headline step with its device time (cold-cache, serialised: compare SHARES)
$CMD > gpurun_out/r02j_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 579 -c 400 --csv --log-file gpurun_out/r02j_launches_wavlm.csv $CMD > gpurun_out/r02j_ncu1.log 2>&1
# 2. --set full of the front end and two encoder layers of the same step (report stays on the box, its raw page comes back)
$CMD > gpurun_out/r02j_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"conv0_tc_kernel|gemm_bf16_tcgen05_2cta_kernel|attention_tc_kernel|layernorm_rows_kernel" -s 549 -c 30 -o /tmp/r02j_prof_wavlm $CMD > gpurun_out/r02j_ncu2.log 2>&1
ncu -i /tmp/r02j_prof_wavlm.ncu-rep --page raw --csv > gpurun_out/r02j_ncu_wavlm_raw.csv 2> /dev/null
# 3. Whisper: attention + log-mel
CMDW="python bench.py --steps 1 --warmup 3 --workload whisper-large-v3 --no-cpu-baseline"
$CMDW > gpurun_out/r02j_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"attention_tc_kernel|logmel_power_kernel" -s 100 -c 3 -o /tmp/r02j_prof_whisper $CMDW > gpurun_out/r02j_ncu3.log 2>&1
ncu -i /tmp/r02j_prof_whisper.ncu-rep --page raw --csv > gpurun_out/r02j_ncu_whisper_raw.csv 2> /dev/null
ls -la gpurun_out/ /tmp/*.ncu-rep
tail -3 gpurun_out/r02j_ncu2.log
