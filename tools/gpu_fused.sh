#!/bin/bash
# scratch script for one gpurun call: the documented runtime switches still give a correct step
set -x
for e in RIFT_B200_STREAMS=0 RIFT_B200_PDL=0 RIFT_B200_CUDA_GRAPH=0 RIFT_B200_WGRAD_GROUP=0 RIFT_B200_FUSE_LN_PLANES=0 RIFT_B200_FUSE_ATTN_PLANES=0 RIFT_B200_ATTN_BWD_TILED=0 RIFT_B200_ATTN_FWD_TILED=0 RIFT_B200_WGRAD_ATOMIC=0 RIFT_B200_SMALL_WGRAD_SIDE=0 RIFT_B200_FUSED=1 RIFT_B200_PDL_LATE=0; do
  echo "== $e"; env $e timeout 300 python -m pytest tests/test_model_gpu.py -m gpu -q -x -k "full_backward_matches_reference_golden or three_policy or baseline_shape_parity_vs_oracle and cfg2-" 2>&1 | tail -1
done
