set -x
nvidia-smi -L
(timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x --timeout 300 -p no:cacheprovider -k "attention" 2>&1 | tail -40) > gpurun_out/r02c_attn_ops.log 2>&1
tail -3 gpurun_out/r02c_attn_ops.log
trace_all() {
  timeout 300 python tools/trace_attn.py
  MODEL=openai/whisper-large-v3 B=32 T=1500 timeout 300 python tools/trace_attn.py
  MODEL=facebook/hubert-xlarge-ls960-ft B=64 T=399 timeout 300 python tools/trace_attn.py
  MODEL=facebook/wav2vec2-xls-r-2b B=64 T=399 timeout 300 python tools/trace_attn.py
  B=16 T=999 timeout 300 python tools/trace_attn.py
}
trace_all > gpurun_out/r02c_trace_v3.log 2>&1
grep "===" gpurun_out/r02c_trace_v3.log
# same-box A/B against the round-1 kernels and the FMA-pipe exp2 variant (development build with the getenv switches)
(SERENC_AB_ARMS=1 python -m interspeech_ser_b200.build) > /dev/null 2>&1
export SERENC_AB_ARMS=1
(SERENC_ATTN_OLD=1 trace_all) > gpurun_out/r02c_trace_old.log 2>&1
grep "===" gpurun_out/r02c_trace_old.log
(SERENC_ATTN_POLY=1 MODEL=openai/whisper-large-v3 B=32 T=1500 timeout 300 python tools/trace_attn.py) > gpurun_out/r02c_trace_poly.log 2>&1
grep "===" gpurun_out/r02c_trace_poly.log
unset SERENC_AB_ARMS
python -m interspeech_ser_b200.build
(timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider 2>&1 | tail -150) > gpurun_out/r02c_pytest.log 2>&1
tail -8 gpurun_out/r02c_pytest.log
(timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02c_bench.json) 2> gpurun_out/r02c_bench.err
tail -c 600 gpurun_out/r02c_bench.err; wc -c gpurun_out/r02c_bench.json
