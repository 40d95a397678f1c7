#!/bin/bash
# The -m gpu suite under every tree (run under gpurun): default environment (the SAH builder), then PTB_BVH=lbvh (Karras
# hierarchy) and PTB_BVH=wide (compressed 8-wide tree), logs under gpurun_out/.
tag=${1:-suite}
(time python -m pytest tests -m gpu -q --durations=5) > gpurun_out/${tag}_default.log 2>&1; tail -6 gpurun_out/${tag}_default.log
(time PTB_BVH=lbvh python -m pytest tests -m gpu -q --durations=5) > gpurun_out/${tag}_lbvh.log 2>&1; tail -6 gpurun_out/${tag}_lbvh.log
(time PTB_BVH=wide python -m pytest tests -m gpu -q --durations=5) > gpurun_out/${tag}_wide.log 2>&1; tail -6 gpurun_out/${tag}_wide.log
# ... and under the 32-byte quantised traversal nodes (compile-time option; build it first: scripts/build_variant.sh qnodes "-DPTB_QNODES=1")
if [ -f build/variants/libptb200_qnodes.so ]; then
  (time PTB200_LIB=build/variants/libptb200_qnodes.so python -m pytest tests -m gpu -q --durations=5) > gpurun_out/${tag}_qnodes.log 2>&1; tail -6 gpurun_out/${tag}_qnodes.log
fi
