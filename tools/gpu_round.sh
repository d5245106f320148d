#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench, then (optionally) ncu launch list + full captures of the top kernels.
# Usage (from the repo root on the GPU box): bash tools/gpu_round.sh [tag] [ncu: 0|1]
TAG=${1:-r02}
NCU=${2:-1}
mkdir -p gpurun_out
set -o pipefail
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_${TAG}.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke_${TAG}.log
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -c 6000 gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
if [ "$NCU" = "1" ]; then
echo "== ncu launch list"
NCU_CMD="python bench.py --steps 1 --warmup 3 --batch 1024 --no-cpu-baseline --no-extras --loss-batch 8192 --profile-range"
timeout 600 $NCU_CMD > gpurun_out/plain_${TAG}.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv $NCU_CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "ncu launches rc=$?"
echo "== ncu full (top kernel)"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_bf16 -s 1 -c 5 -o gpurun_out/prof_gemm_${TAG} $NCU_CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "ncu full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"attention_fwd|clip_loss_fwd|clip_loss_grad|layernorm|row_stats|patch_embed|pool_head" -s 0 -c 10 -o gpurun_out/prof_attn_loss_${TAG} $NCU_CMD > gpurun_out/ncu_full2_${TAG}.log 2>&1
echo "ncu full (attention/loss) rc=$?"
fi
ls -la gpurun_out | tail -8
