#!/bin/bash
# fused mirror (quaternion track kernel with the mirror epilogue): parity, sanitizer pass, timing against the two-kernel path
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ik.py tests/test_abi.py -x -q -m gpu 2>&1 | tail -12 | cut -c1-250
rm -f gpurun_out/r2_mirror.jsonl
export PMB_EXPERIMENT=1
for cfg in "PMB_MIRROR_FUSED=0" "PMB_MIRROR_FUSED=1" "PMB_MIRROR_FUSED=1 PMB_QT_PIPE=0" "PMB_MIRROR_FUSED=1 PMB_QT_PIPE=1"; do
  for wl in fk_1m_x_22 fk_4m_x_52 fk_4m_x_65; do
    env $cfg timeout 300 python bench.py --kernel-only --steps 20 --warmup 3 --op mirror_all --workload $wl | sed "s/^{/{\"cfg\": \"$cfg\", /" >> gpurun_out/r2_mirror.jsonl
  done
done
cut -c1-330 gpurun_out/r2_mirror.jsonl
