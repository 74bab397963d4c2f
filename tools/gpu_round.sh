#!/bin/bash
# Runs on the GPU box under gpurun: staged so that a hang in one stage cannot hide the others.
# usage: tools/gpu_round.sh [stages...]   (default: all)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
stages="${@:-info kernels probe umma convbwd model_fp32 model smoke bench_small bench}"
run() { # name timeout cmd...
  local name=$1 t=$2; shift 2
  echo "=== $name ===" | tee -a $OUT/summary.txt
  timeout $t "$@" > $OUT/$name.log 2>&1
  local rc=$?
  echo "$name rc=$rc" | tee -a $OUT/summary.txt
  tail -n 15 $OUT/$name.log | tee -a $OUT/summary.txt
}
for s in $stages; do
case $s in
  info) run info 60 bash -c 'nvidia-smi; nproc; free -g | head -2; python -c "import torch;print(torch.__version__, torch.cuda.get_device_name(0))"' ;;
  kernels) run kernels 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "not umma and not probe and not conv_backward" -p no:cacheprovider ;;
  probe) run probe 120 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "probe" -p no:cacheprovider ;;
  umma) run umma 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "umma" -p no:cacheprovider ;;
  convbwd) run convbwd 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv_backward" -p no:cacheprovider ;;
  model_fp32) PVG_PRECISION=fp32 run model_fp32 900 python -m pytest tests/test_model_gpu.py -q -m gpu -p no:cacheprovider ;;
  model) run model 900 python -m pytest tests/test_model_gpu.py -q -m gpu -p no:cacheprovider ;;
  smoke) run smoke 300 python -c "import __graft_entry__ as g; g.smoke()" ;;
  bench_small) run bench_small 600 python bench.py --workload bair64_b2_t4 --steps 3 --warmup 3 --no-cpu-baseline ;;
  ncu_stem) run ncu_stem 300 ncu --set full --clock-control none --import-source on -k regex:conv_stem3 -s 2 -c 1 -f -o $OUT/r02_stem python tools/stem_bench.py 120; python tools/ncu_extract.py $OUT/r02_stem.ncu-rep ;;
  stem) run stem 300 python tools/stem_bench.py; run stem_t 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "stem or producer_planes" -p no:cacheprovider ;;
  layers) run layers 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline --no-secondary --no-parity-check --layer-table $OUT/r02_layer_table.md ;;
  ncu_epi) run ncu_epi 300 ncu --set full --clock-control none --import-source on -k regex:conv_h3 -s 3 -c 1 -f -o $OUT/r02_epi python tools/conv_micro.py 120 64 128 128 128 planes; python tools/ncu_extract.py $OUT/r02_epi.ncu-rep ;;
  bn_micro) run bn_micro 300 python tools/bn_micro.py ;;      # A/B against another build: PVG_LIB=/path/to/libpvg_b200_other.so
  ncu_final) for spec in "vgg2_1_planes 120 64 128 128 128 planes" "vgg3_2_planes 120 256 256 64 64 planes" "dec_res_sums 8 128 128 64 64 sums" "enc_res_sums 128 64 64 32 32 sums" "vgg1_2_relu 120 64 64 256 256 bias+relu"; do set -- $spec; run ncu_$1 200 ncu --set full --clock-control none --import-source on -k regex:conv_h3 -s 3 -c 1 -f -o $OUT/r02f_$1 python tools/conv_micro.py $2 $3 $4 $5 $6 $7; python tools/ncu_extract.py $OUT/r02f_$1.ncu-rep; done ;;
  micro) run micro 600 python tools/conv_micro.py ;;
  micro_lstm) run micro_lstm 300 python tools/conv_micro.py lstm ;;
  benchq) run benchq 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline --no-secondary --no-parity-check ;;
  benchq_notags) PVG_NO_AMAX_TAGS=1 run benchq_notags 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline --no-secondary --no-parity-check ;;
  bench) run bench 600 python bench.py --steps 3 --warmup 3 ;;
  diag_fp32) run diag_fp32 600 python tools/grad_diag.py full_bair_feedback fp32 ;;
  diag_tf32x3) run diag_tf32x3 600 python tools/grad_diag.py full_bair_feedback tf32x3 ;;
  allkernels) run allkernels 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider ;;
  bench2) run bench2 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline ;;
  bench_nograph) run bench_nograph 600 python bench.py --steps 3 --warmup 2 --no-graph --no-cpu-baseline ;;
  convbench) run convbench 600 python tools/conv_bench.py tf32x3 5 ;;
  convbench1) run convbench1 600 python tools/conv_bench.py tf32 5 ;;
  ncu_conv) run ncu_conv 900 ncu --set full --clock-control none --import-source on -k regex:conv_umma_kernel -c 10 -f -o $OUT/prof_conv python tools/conv_bench.py tf32x3 1 ;;
  allkernels2) run allkernels2 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider ;;
  ncu_traffic) run ncu_traffic 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_umma_kernel -s 800 -c 790 --csv --log-file $OUT/conv_traffic.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph ;;
  pairs_k) PVG_2CTA=1 run pairs_k 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "umma_forward or conv_backward" -p no:cacheprovider ;;
  pairs_b) PVG_2CTA=1 run pairs_b 300 python tools/conv_bench.py tf32x3 5 ;;
  pairs_b1) PVG_2CTA=1 run pairs_b1 300 python tools/conv_bench.py tf32 5 ;;
  directb) run directb 300 python tools/direct_bench.py ;;
  directb0) PVG_NO_DIRECT=1 run directb0 300 python tools/direct_bench.py ;;
  wgradb) run wgradb 300 python tools/wgrad_bench.py tf32x3 ;;
  rollout) run rollout 300 python tools/rollout_bench.py 64 100 tf32x3 ;;
  stepprof) run stepprof 400 python tools/step_profile.py tf32x3 gpurun_out/step_profile.json ;;
  flaky) run flaky 300 python tools/flaky_probe.py ;;
  convbench16) PVG_KC=16 run convbench16 600 python tools/conv_bench.py tf32x3 5 ;;
  kernels16) PVG_KC=16 run kernels16 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "umma or conv_backward" -p no:cacheprovider ;;
  bench_ref) run bench_ref 900 python bench.py --impl reference --steps 1 --warmup 0 ;;
  ncu_list) run ncu_list 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60000 --csv --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph ;;
  lb_heads) run lb_heads 300 python tools/layer_bench.py tf32x3 heads ;;
  lb_enc) run lb_enc 300 python tools/layer_bench.py tf32x3 enc ;;
  lb_vgg) run lb_vgg 300 python tools/layer_bench.py tf32x3 vgg ;;
  lb_model) run lb_model 300 python tools/layer_bench.py tf32x3 model ;;
  lb_all) run lb_all 600 python tools/layer_bench.py tf32x3 all ;;
  t_direct) run t_direct 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv_simt or conv_backward" -p no:cacheprovider ;;
  t_conv) run t_conv 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv" -p no:cacheprovider ;;
  diag_pre) for c in "tf32 tf32 tf32" "fp16 tf32 tf32" "tf32 bf16 tf32" "fp16 bf16 bf16"; do set -- $c; PVG_CORR=$1 PVG_DGRAD_CORR=$2 PVG_WGRAD_CORR=$3 run diag_pre_$1_$2_$3 300 python -m pytest tests/test_model_gpu.py -q -m gpu -k "pretrain_bair" -p no:cacheprovider; grep pretrain_bair:cond $OUT/model_errors.jsonl | tail -1 | cut -c1-400 | tee -a $OUT/summary.txt; done ;;
  bench_fwdtf32) PVG_CORR=tf32 run bench_fwdtf32 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline ;;
  bench_alltf32) PVG_CORR=tf32 PVG_DGRAD_CORR=tf32 PVG_WGRAD_CORR=tf32 run bench_alltf32 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline ;;
  lb_pairs) PVG_2CTA=1 run lb_pairs 300 python tools/layer_bench.py tf32x3 vgg; PVG_2CTA=1 run lb_pairs_model 300 python tools/layer_bench.py tf32x3 model ;;
  lb_nopairs) PVG_2CTA=0 run lb_nopairs 300 python tools/layer_bench.py tf32x3 vgg ;;
  lb_tf32) run lb_tf32 300 python tools/layer_bench.py tf32 vgg ;;
  ncu_vgg3) run ncu_vgg3 600 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 2 -c 1 -f -o $OUT/prof_vgg3 python tools/one_conv.py 120 256 256 64 64 ;;
  ncu_vgg1) run ncu_vgg1 600 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 2 -c 1 -f -o $OUT/prof_vgg1 python tools/one_conv.py 120 64 64 256 256 ;;
  ncu_vgg3_tf32) PVG_PRECISION=tf32 run ncu_vgg3_tf32 600 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 2 -c 1 -f -o $OUT/prof_vgg3_tf32 python tools/one_conv.py 120 256 256 64 64 ;;
  tile_model) PVG_2CTA=0 run tile_model0 300 python tools/tile_model.py tf32x3; PVG_2CTA=1 run tile_model1 300 python tools/tile_model.py tf32x3; PVG_2CTA=0 run tile_model0_tf32 300 python tools/tile_model.py tf32; PVG_2CTA=0 run tile_model0_64 300 python tools/tile_model.py tf32x3 64; PVG_2CTA=0 PVG_CORR=tf32 run tile_model0_c3 300 python tools/tile_model.py tf32x3 ;;
  head_stress) run head_stress 300 python tools/head_stress.py 60 ;;
  head_san) run head_race 600 compute-sanitizer --tool racecheck --print-limit 5 python tools/head_stress.py 1; run head_init 600 compute-sanitizer --tool initcheck --print-limit 5 python tools/head_stress.py 1; run head_mem 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/head_stress.py 1 ;;
  san_tests) PYTORCH_NO_CUDA_MEMORY_CACHING=1 run san_tests 500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "conv_simt or probe or umma_forward_tf32x3 or conv_backward" -p no:cacheprovider ;;
  graphed_rollout) run t_graphed_rollout 200 python -m pytest tests/test_model_gpu.py -q -m gpu -k "graphed_rollout" -p no:cacheprovider; run rollout_b1 100 python tools/rollout_bench.py 1 100 tf32x3; run rollout_b1_graph 100 python tools/rollout_bench.py 1 100 tf32x3 graph ;;
  desc_probe) run desc_probe 300 bash -c 'cd tools/probes && nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../playablevideogeneration_b200/csrc -o /tmp/umma_desc_probe umma_desc_probe.cu && /tmp/umma_desc_probe' ;;
  persist_t) PVG_PERSISTENT=1 PVG_2CTA=0 run persist_t 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv" -p no:cacheprovider ;;
  persist_b) PVG_PERSISTENT=1 PVG_2CTA=0 run persist_b 300 python tools/tile_model.py tf32x3; PVG_PERSISTENT=1 PVG_2CTA=0 run persist_b64 300 python tools/tile_model.py tf32x3 64 ;;
  persist2_t) PVG_PERSISTENT=1 run persist2_t 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv" -p no:cacheprovider ;;
  persist2_b) PVG_PERSISTENT=1 PVG_2CTA=1 run persist2_b 300 python tools/tile_model.py tf32x3; PVG_PERSISTENT=1 run persist2_lb 300 python tools/layer_bench.py tf32x3 vgg ;;
  corr_diag) run corr_diag 300 python tools/corr_diag.py ;;
  h3_ab) for rep in 1 2; do PVG_H3_LEGACY=1 PVG_PERSISTENT=0 run h3ab_legacy_$rep 200 python tools/tile_model.py tf32x3; for hp in "0 0" "1 0" "0 1" "1 1"; do set -- $hp; PVG_H3_HALO=$1 PVG_H3_PAIR=$2 run h3ab_$1$2_$rep 200 python tools/tile_model.py tf32x3; done; nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv | tee -a $OUT/summary.txt; done ;;
  h3_ab2) PVG_H3_LEGACY=1 PVG_PERSISTENT=0 run h3ab_legacy 200 python tools/tile_model.py tf32x3; for hpd in "0 0 0" "1 0 0" "1 0 1" "0 1 0" "1 1 0" "1 1 1"; do set -- $hpd; PVG_H3_HALO=$1 PVG_H3_PAIR=$2 PVG_H3_DBG=$3 run h3ab_$1$2$3 200 python tools/tile_model.py tf32x3; done ;;
  ncu_h3) PVG_H3_HALO=1 PVG_H3_PAIR=0 run ncu_h3_10 400 ncu --set full --clock-control none --import-source on -k regex:conv_h3 -s 2 -c 1 -f -o $OUT/prof_h3_10 python tools/one_conv.py 120 256 128 64 64; PVG_H3_HALO=1 PVG_H3_PAIR=1 run ncu_h3_11 400 ncu --set full --clock-control none --import-source on -k regex:conv_h3 -s 2 -c 1 -f -o $OUT/prof_h3_11 python tools/one_conv.py 120 256 128 64 64; PVG_H3_LEGACY=1 PVG_PERSISTENT=0 run ncu_h3_legacy 400 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 2 -c 1 -f -o $OUT/prof_h3_legacy python tools/one_conv.py 120 256 128 64 64 ;;
  h3_ab3) PVG_H3_LEGACY=1 PVG_PERSISTENT=0 run h3ab_legacy 200 python tools/tile_model.py tf32x3; for hp in "0 0" "1 0" "0 1" "1 1"; do set -- $hp; PVG_H3_HALO=$1 PVG_H3_PAIR=$2 run h3_t_$1$2 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "umma_forward_tf32x3 and h3" -p no:cacheprovider; PVG_H3_HALO=$1 PVG_H3_PAIR=$2 run h3ab_$1$2 200 python tools/tile_model.py tf32x3; done; PVG_H3_HALO=1 PVG_H3_PAIR=1 run h3ab64_11 200 python tools/tile_model.py tf32x3 64; PVG_H3_HALO=1 PVG_H3_PAIR=0 run h3ab64_10 200 python tools/tile_model.py tf32x3 64 ;;
  h3_ab4) for hpd in "1 0 0" "1 0 1" "0 0 0"; do set -- $hpd; PVG_H3_HALO=$1 PVG_H3_PAIR=$2 PVG_H3_DBG=$3 run h3ab_$1$2$3 200 python tools/tile_model.py tf32x3; PVG_H3_HALO=$1 PVG_H3_PAIR=$2 PVG_H3_DBG=$3 run h3ab64_$1$2$3 200 python tools/tile_model.py tf32x3 64; done ;;
  ncu_h3b) PVG_H3_HALO=1 PVG_H3_PAIR=0 run ncu_h3b_10 400 ncu --set full --clock-control none --import-source on -k regex:conv_h3 -s 2 -c 1 -f -o $OUT/prof_h3b_10 python tools/one_conv.py 120 256 64 64 64; PVG_H3_HALO=0 PVG_H3_PAIR=0 run ncu_h3b_00 400 ncu --set full --clock-control none --import-source on -k regex:conv_h3 -s 2 -c 1 -f -o $OUT/prof_h3b_00 python tools/one_conv.py 120 256 64 64 64 ;;
  ncu_wgrad) run ncu_wgrad_a 400 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_umma -s 2 -c 1 -f -o $OUT/prof_wgrad_a python tools/one_wgrad.py 8 64 32 256 256; run ncu_wgrad_b 400 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_umma -s 2 -c 1 -f -o $OUT/prof_wgrad_b python tools/one_wgrad.py 8 128 128 64 64; run wgradb 300 python tools/wgrad_bench.py tf32x3 ;;
  ncu_list2) run ncu_list2 1700 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60000 --csv --log-file $OUT/launches_r02.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph --no-gpu-eager-baseline --no-secondary --no-parity-check ;;
  ncu_r02) for spec in "vgg3_2 120 256 256 64 64" "enc_res0 128 16 16 128 128" "dec_up2 8 64 32 256 256" "lstm2_gates 8 288 512 32 32" "vgg1_2 120 64 64 256 256"; do set -- $spec; run ncu_$1 300 ncu --set full --clock-control none --import-source on -k regex:conv_h3 -s 2 -c 1 -f -o $OUT/r02_$1 python tools/one_conv.py $2 $3 $4 $5 $6; python tools/ncu_extract.py $OUT/r02_$1.ncu-rep; done; for spec in "wgrad_dec_up2 8 64 32 256 256" "wgrad_dec_res0 8 128 128 64 64"; do set -- $spec; run ncu_$1 300 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_umma -s 2 -c 1 -f -o $OUT/r02_$1 python tools/one_wgrad.py $2 $3 $4 $5 $6; python tools/ncu_extract.py $OUT/r02_$1.ncu-rep; done ;;
  h3_matrix) for hp in "0 0" "1 0" "0 1" "1 1"; do set -- $hp; PVG_H3_HALO=$1 PVG_H3_PAIR=$2 run h3_t_$1$2 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "umma_forward_tf32x3 and h3" -p no:cacheprovider; PVG_H3_HALO=$1 PVG_H3_PAIR=$2 run h3_tm_$1$2 200 python tools/tile_model.py tf32x3; PVG_H3_HALO=$1 PVG_H3_PAIR=$2 run h3_tm64_$1$2 200 python tools/tile_model.py tf32x3 64; done ;;
  t_umma) run t_umma 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "umma or conv_backward" -p no:cacheprovider ;;
  tm_h3) PVG_2CTA=0 PVG_PERSISTENT=0 run tm_h3 300 python tools/tile_model.py tf32x3; PVG_2CTA=0 PVG_PERSISTENT=0 run tm_h3_64 300 python tools/tile_model.py tf32x3 64; PVG_2CTA=0 PVG_PERSISTENT=0 PVG_CORR=fp16 run tm_fp16 300 python tools/tile_model.py tf32x3 ;;
esac
done
cp $OUT/summary.txt $OUT/summary_$(date +%s).txt 2>/dev/null
true
