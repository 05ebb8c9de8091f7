set -x
rm -f gpurun_out/parity_log.jsonl
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8
grep fused gpurun_out/parity_log.jsonl
