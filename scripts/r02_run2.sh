set -x
nvidia-smi -L
(timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x --timeout 300 -p no:cacheprovider -k "attention" 2>&1 | tail -40) > gpurun_out/r02b_attn_ops.log 2>&1
tail -3 gpurun_out/r02b_attn_ops.log
(timeout 300 python tools/trace_attn.py; MODEL=openai/whisper-large-v3 B=32 T=1500 timeout 300 python tools/trace_attn.py) > gpurun_out/r02b_trace_split.log 2>&1
grep "===" gpurun_out/r02b_trace_split.log
# same-box A/B against the first-generation kernel (development build with the getenv switches)
(SERENC_AB_ARMS=1 python -m interspeech_ser_b200.build && export SERENC_AB_ARMS=1 SERENC_ATTN_GEN1=1 && timeout 300 python tools/trace_attn.py && MODEL=openai/whisper-large-v3 B=32 T=1500 timeout 300 python tools/trace_attn.py) > gpurun_out/r02b_trace_gen1.log 2>&1
grep "===" gpurun_out/r02b_trace_gen1.log
python -m interspeech_ser_b200.build
(timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider 2>&1 | tail -150) > gpurun_out/r02b_pytest.log 2>&1
tail -8 gpurun_out/r02b_pytest.log
(timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02b_bench.json) 2> gpurun_out/r02b_bench.err
tail -c 1000 gpurun_out/r02b_bench.err; wc -c gpurun_out/r02b_bench.json
