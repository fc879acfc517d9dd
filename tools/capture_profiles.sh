#!/bin/bash
# One gpurun call that refreshes the ncu artefacts tools/summarize_profiles.py turns into profiles/<tag>_*.
# usage (on the GPU box): bash tools/capture_profiles.sh r2
# The `--set full` reports are exported to CSV on the box and deleted: gpurun copies back at most 64 MiB.
tag=${1:-r2}
mkdir -p gpurun_out
NCU="ncu --profile-from-start off --clock-control none"
list() {   # name, script + args
  local name=$1; shift
  timeout -s KILL 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_${tag}_${name}.csv python "$@" > gpurun_out/cap_${name}.log 2>&1
  echo "launch list ${name} rc=$?"
}
full() {   # name, kernel regex, skip, count, script + args
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout -s KILL 900 $NCU --set full -k regex:"${rx}" -s ${skip} -c ${cnt} -f -o gpurun_out/prof_${name}_${tag} python "$@" > gpurun_out/cap_${name}.log 2>&1
  echo "full capture ${name} rc=$?"
  ncu -i gpurun_out/prof_${name}_${tag}.ncu-rep --page raw --csv > gpurun_out/prof_${name}_${tag}.csv 2>/dev/null
  rm -f gpurun_out/prof_${name}_${tag}.ncu-rep
}
list dit tools/profile_step.py
list dit_fp8 tools/profile_step.py --fp8
list vae tools/profile_vae.py
# --set full: 3 blocks are enough (replays are slow); launches of block 0 are skipped
full attn attention 2 2 tools/profile_step.py --layers 3
full gemm gemm_bf16_kernel 8 8 tools/profile_step.py --layers 3
full gemm8 gemm_bf16_kernel 8 6 tools/profile_step.py --layers 3 --fp8
full rows "norm_modulate|qkv_head_scatter|headnorm_rope" 6 4 tools/profile_step.py --layers 3
full conv conv3d 36 8 tools/profile_vae.py
full vaerow norm_act_pad 3 2 tools/profile_vae.py
rm -f gpurun_out/cap_*.log
ls -la gpurun_out | tail -20
du -sh gpurun_out
