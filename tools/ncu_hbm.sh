#!/bin/bash
# ncu --set full capture of each HBM-bound kernel at one large size (second launch of each; one GPU).  The
# .ncu-rep files stay on the box (/tmp); the summaries (tools/ncu_summary.py, tools/ncu_lines.py) come back.
#   bash tools/ncu_hbm.sh <tag> [rays] [samples] [kernel regexes...]
TAG=${1:-r1b}; N=${2:-65536}; S=${3:-128}; shift 3
KERNELS=${@:-"composite_fwd composite_bwd sample_pdf_fast sample_pdf_mu_sigma_fast dp_loss_fwd_fast dp_loss_bwd_fast first_cycle encode_img"}
declare -A OBJ=( [composite_fwd]=composite [composite_bwd]=composite [sample_pdf_fast]=sampler [sample_pdf_mu_sigma_fast]=sampler
                 [dp_loss_fwd_fast]=dploss [dp_loss_bwd_fast]=dploss [first_cycle]=sampler [encode_img]=mlp_tc
                 [composite_dd_fwd]=composite [composite_dd_bwd]=composite )
mkdir -p /tmp/ncu_$TAG gpurun_out
for k in $KERNELS; do
  ncu --set full --clock-control none --import-source on -k regex:${k}_kernel -s 1 -c 1 -f \
      -o /tmp/ncu_$TAG/${k} python tools/hbm_kernels_once.py $N $S > gpurun_out/${TAG}_ncu_${k}.log 2>&1
  if [ -f /tmp/ncu_$TAG/${k}.ncu-rep ]; then
    python tools/ncu_summary.py /tmp/ncu_$TAG/${k}.ncu-rep > gpurun_out/${TAG}_ncu_${k}.md 2>> gpurun_out/${TAG}_ncu_${k}.log
    echo -e "\n## executed instructions and stall samples by source line\n\n\`\`\`" >> gpurun_out/${TAG}_ncu_${k}.md
    python tools/ncu_lines.py /tmp/ncu_$TAG/${k}.ncu-rep ddnerf_b200/build/${OBJ[$k]}.o 25 >> gpurun_out/${TAG}_ncu_${k}.md 2>> gpurun_out/${TAG}_ncu_${k}.log
    echo '```' >> gpurun_out/${TAG}_ncu_${k}.md
    echo "$k ok"
  else
    echo "$k: no report"; tail -2 gpurun_out/${TAG}_ncu_${k}.log
  fi
done
