#!/bin/bash
# role-isolation experiments on the conv kernel (see MPNN_TUNE_DBG in stencil_umma.cu)
for dbg in 0 1 2 4 8 3 5 6 7 15; do
  echo "== DBG=$dbg"; MPNN_TUNE_DBG=$dbg python scratch/mb_conv.py gemm1 2>&1 | grep gemm
done
echo "== HALO8"; MPNN_TUNE_HALO8=1 python scratch/mb_conv.py gemm1 2>&1 | grep gemm
for ps in 1 2; do echo "== PER_SM=$ps DBG=1"; MPNN_TUNE_PER_SM=$ps MPNN_TUNE_DBG=1 python scratch/mb_conv.py gemm1 | grep gemm;
 echo "== PER_SM=$ps DBG=6"; MPNN_TUNE_PER_SM=$ps MPNN_TUNE_DBG=6 python scratch/mb_conv.py gemm1 | grep gemm;
 echo "== PER_SM=$ps DBG=15"; MPNN_TUNE_PER_SM=$ps MPNN_TUNE_DBG=15 python scratch/mb_conv.py gemm1 | grep gemm; done
