#!/bin/bash
# ncu --set full capture of each HBM-bound kernel at one large size (second launch of each; one GPU).
#   bash tools/ncu_hbm.sh <tag> [rays] [samples] [kernel regexes...]
TAG=${1:-r1b}; N=${2:-65536}; S=${3:-128}; shift 3
KERNELS=${@:-"composite_fwd composite_bwd sample_pdf_fast sample_pdf_mu_sigma_fast dp_loss_fwd_fast dp_loss_bwd_fast first_cycle"}
for k in $KERNELS; do
  ncu --set full --clock-control none --import-source on -k regex:${k}_kernel -s 1 -c 1 -f \
      -o gpurun_out/${TAG}_ncu_${k} python tools/hbm_kernels_once.py $N $S > gpurun_out/${TAG}_ncu_${k}.log 2>&1
  tail -1 gpurun_out/${TAG}_ncu_${k}.log
done
