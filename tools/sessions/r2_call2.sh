#!/bin/bash
# Round 2, call 2: new MSDeformAttn kernel families (run / wide) -- parity, op benchmark; plus the prep-work A/B.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "=== op tests"
timeout 1200 python -m pytest tests/test_msda_gpu.py -q --maxfail=12 --tb=line 2>&1 | tail -30
echo "=== opbench"
timeout 900 python tools/opbench.py --cases c2_enc_init,c2_enc_smooth,c2_enc_model,c2_enc_uniform,c2_dec,c2_enc_init_n2,c2_dec_n2 \
   --variants 0,100,101,110,20 --bwd-variants 0,100,101,110,20 --iters 20 --out gpurun_out/r2_opbench_v1.json 2>&1 | cut -c1-230
echo "=== tcgen05 GEMM (own timeout: a wrong barrier would hang)"
timeout 300 python -m pytest tests/test_tf32_gemm_gpu.py -q --maxfail=6 --tb=line 2>&1 | tail -14
timeout 300 python tools/gemm_bench.py --iters 10 2>&1 | cut -c1-400 | tail -12
echo "=== prep tests + fused loss / refine tests"
timeout 900 python -m pytest tests/test_fused_loss_gpu.py tests/test_fused_bn_gpu.py tests/test_fused_norm_gpu.py tests/test_train_step_gpu.py -q --tb=line 2>&1 | tail -16
run() {
  local name=$1; shift
  env "$@" timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_r2_${name}.err | tee gpurun_out/bench_r2_${name}.json | cut -c1-170
}
run all_on      TFB200_FUSED_BN=1 TFB200_GATHER_GRADS=1 TFB200_LN_SEEDED=1
run no_fused_bn TFB200_FUSED_BN=0 TFB200_GATHER_GRADS=1 TFB200_LN_SEEDED=1
run no_gather   TFB200_FUSED_BN=1 TFB200_GATHER_GRADS=0 TFB200_LN_SEEDED=1
run no_seeded   TFB200_FUSED_BN=1 TFB200_GATHER_GRADS=1 TFB200_LN_SEEDED=0
run fused_loss  TFB200_FUSED_LOSS=1
run tcgen05     TFB200_TCGEN05_LINEAR=1
run fused_prep  TFB200_FUSED_PREP=1
run all_new     TFB200_FUSED_LOSS=1 TFB200_TCGEN05_LINEAR=1 TFB200_FUSED_PREP=1
echo "=== reference classes on the same GPU: (a) on our extension (zero-edit drop-in), (b) on the reference's own CUDA kernels"
timeout 600 python tools/ref_gpu_bench.py --msda ours --check --out gpurun_out/r2_ref_gpu_ours.json 2>&1 | tail -2 | cut -c1-400
timeout 600 python tools/ref_gpu_bench.py --msda refcuda --out gpurun_out/r2_ref_gpu_refcuda.json 2>&1 | tail -2 | cut -c1-400
echo "=== model parity (incl. the new full-size cases)"
timeout 1500 python -m pytest tests/test_model_parity_gpu.py tests/test_tracker_gpu.py -q -s --tb=line 2>&1 | grep -v "^$" | tail -60
echo "=== ncu launch list of the step graph (kernel nodes)"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 14000 --csv \
  --log-file gpurun_out/r2_launches_head.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/r2_launches_head.csv
gzip -f gpurun_out/r2_launches_head.csv
