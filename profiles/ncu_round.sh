#!/bin/bash
# Round profile set (run under gpurun, 1 GPU): launch list of one step + full captures of the
# dominant tensor kernel (ViT GEMMs, CTA-pair), the tcgen05 attention kernel and the HBM-bound
# FeatureFusionBlock kernel (bilinear x2 + skip add).  Usage: profiles/ncu_round.sh <tag>
TAG=${1:-r01b}
mkdir -p gpurun_out
COMMON="--clock-control none --profile-from-start off"
timeout 600 ncu --metrics gpu__time_duration.sum $COMMON --csv --log-file gpurun_out/${TAG}_launches.csv python profiles/ncu_target.py > gpurun_out/${TAG}_ncu_a.log 2>&1
timeout 600 ncu --set full --import-source on $COMMON -k regex:conv_gemm -s 53 -c 4 -o gpurun_out/${TAG}_vit_gemm python profiles/ncu_target.py > gpurun_out/${TAG}_ncu_b.log 2>&1
timeout 600 ncu --set full --import-source on $COMMON -k regex:attention_tc -c 1 -o gpurun_out/${TAG}_attention python profiles/ncu_target.py > gpurun_out/${TAG}_ncu_c.log 2>&1
timeout 600 ncu --set full --import-source on $COMMON -k regex:upsample2x -s 2 -c 3 -o gpurun_out/${TAG}_upsample python profiles/ncu_target.py > gpurun_out/${TAG}_ncu_d.log 2>&1
timeout 600 ncu --set full $COMMON -k regex:groupnorm_apply -s 2 -c 2 -o gpurun_out/${TAG}_gn_apply python profiles/ncu_target.py > gpurun_out/${TAG}_ncu_e.log 2>&1
timeout 600 ncu --set full --import-source on $COMMON -k regex:conv_gemm -s 129 -c 1 -o gpurun_out/${TAG}_head_conv python profiles/ncu_target.py > gpurun_out/${TAG}_ncu_f.log 2>&1
for f in vit_gemm attention upsample gn_apply head_conv; do
  ncu -i gpurun_out/${TAG}_${f}.ncu-rep --page raw --csv > gpurun_out/${TAG}_${f}_ncu_full_metrics.csv 2>/dev/null
done
ls -la gpurun_out | grep ${TAG}
