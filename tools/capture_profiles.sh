#!/bin/bash
# One gpurun call that refreshes the ncu artefacts tools/summarize_profiles.py turns into profiles/<tag>_*.
# usage (on the GPU box): bash tools/capture_profiles.sh r1c
tag=${1:-r1}
mkdir -p gpurun_out
NCU="ncu --profile-from-start off --clock-control none"
timeout -s KILL 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_${tag}_dit.csv python tools/profile_step.py > gpurun_out/cap_dit.log 2>&1
echo "dit launch list rc=$?"
timeout -s KILL 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_${tag}_vae.csv python tools/profile_vae.py > gpurun_out/cap_vae.log 2>&1
echo "vae launch list rc=$?"
# --set full captures: 2 blocks are enough (replays are slow); launches of block 0 are skipped
timeout -s KILL 600 $NCU --set full --import-source on -k regex:attention -s 2 -c 2 -f -o gpurun_out/prof_attn_${tag} python tools/profile_step.py --layers 3 > gpurun_out/cap_attn.log 2>&1
echo "attention capture rc=$?"
timeout -s KILL 600 $NCU --set full --import-source on -k regex:gemm_bf16_kernel -s 8 -c 8 -f -o gpurun_out/prof_gemm_${tag} python tools/profile_step.py --layers 3 > gpurun_out/cap_gemm.log 2>&1
echo "gemm capture rc=$?"
timeout -s KILL 600 $NCU --set full -k regex:"norm_modulate|qkv_head_scatter|headnorm_rope" -s 6 -c 4 -f -o gpurun_out/prof_rows_${tag} python tools/profile_step.py --layers 3 > gpurun_out/cap_rows.log 2>&1
echo "row kernels capture rc=$?"
timeout -s KILL 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_${tag}_dit_fp8.csv python tools/profile_step.py --fp8 > gpurun_out/cap_dit8.log 2>&1
echo "dit fp8 launch list rc=$?"
timeout -s KILL 600 $NCU --set full --import-source on -k regex:gemm_bf16_kernel -s 8 -c 6 -f -o gpurun_out/prof_gemm8_${tag} python tools/profile_step.py --layers 3 --fp8 > gpurun_out/cap_gemm8.log 2>&1
echo "fp8 gemm capture rc=$?"
timeout -s KILL 600 $NCU --set full --import-source on -k regex:conv3d -s 36 -c 8 -f -o gpurun_out/prof_conv_${tag} python tools/profile_vae.py > gpurun_out/cap_conv.log 2>&1
echo "conv capture rc=$?"
timeout -s KILL 600 $NCU --set full -k regex:norm_act_pad -s 30 -c 2 -f -o gpurun_out/prof_vaerow_${tag} python tools/profile_vae.py > gpurun_out/cap_vaerow.log 2>&1
echo "vae row capture rc=$?"
ls -la gpurun_out | tail -20
