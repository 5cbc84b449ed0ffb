# final state under the alternate paths (every optional fast path off; tap-gather only; fp32 FFMA only)
mkdir -p gpurun_out
SKIP='not tap_pass and not channels_on and not sixteen and not cfg4_shard and not tensor_core_path'
echo "== NASR_TMA_STORE=0 NASR_PDL=0 NASR_ZEROCOPY=0 NASR_STREAM_GRAPH=0 NASR_HOST_PIPE=0"
NASR_TMA_STORE=0 NASR_PDL=0 NASR_ZEROCOPY=0 NASR_STREAM_GRAPH=0 NASR_HOST_PIPE=0 python -m pytest tests -m gpu -q 2>&1 | tail -3
echo "== NASR_PATH=tc"
NASR_PATH=tc python -m pytest tests -m gpu -q -k "$SKIP" 2>&1 | tail -3
echo "== NASR_PATH=fp32"
NASR_PATH=fp32 python -m pytest tests -m gpu -q -k "$SKIP and not fp16_range" 2>&1 | tail -3
echo "== NASR_LOWER=0"
NASR_LOWER=0 python -m pytest tests -m gpu -q -k "not sixteen" 2>&1 | tail -3
