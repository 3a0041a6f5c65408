set -x
(SERENC_AB_ARMS=1 python -m interspeech_ser_b200.build) > /dev/null 2>&1
export SERENC_AB_ARMS=1 NOTRACE=1
for v in 0 1 2 3 4 0; do
  echo "##### variant $v"
  SERENC_ATTN_VARIANT=$v timeout 300 python tools/trace_attn.py 2>&1 | grep "==="
  SERENC_ATTN_VARIANT=$v MODEL=openai/whisper-large-v3 B=32 T=1500 timeout 300 python tools/trace_attn.py 2>&1 | grep "==="
  SERENC_ATTN_VARIANT=$v B=16 T=999 timeout 300 python tools/trace_attn.py 2>&1 | grep "==="
  if [ $v -le 1 ]; then
    SERENC_ATTN_VARIANT=$v MODEL=facebook/hubert-xlarge-ls960-ft B=64 T=399 timeout 300 python tools/trace_attn.py 2>&1 | grep "==="
    SERENC_ATTN_VARIANT=$v MODEL=facebook/wav2vec2-xls-r-2b B=64 T=399 timeout 300 python tools/trace_attn.py 2>&1 | grep "==="
  fi
done > gpurun_out/r02d_attn_variants.log 2>&1
cat gpurun_out/r02d_attn_variants.log
