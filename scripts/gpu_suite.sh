#!/bin/bash
# The -m gpu suite under both trees (run under gpurun): default environment, then PTB_BVH=wide / binary (whichever is not
# the default), logs under gpurun_out/.
tag=${1:-suite}
(time python -m pytest tests -m gpu -q --durations=5) > gpurun_out/${tag}_default.log 2>&1; tail -6 gpurun_out/${tag}_default.log
(time PTB_BVH=wide python -m pytest tests -m gpu -q --durations=5) > gpurun_out/${tag}_wide.log 2>&1; tail -6 gpurun_out/${tag}_wide.log
# ... and under the 32-byte quantised traversal nodes (compile-time option; build it first: scripts/build_variant.sh qnodes "-DPTB_QNODES=1")
if [ -f build/variants/libptb200_qnodes.so ]; then
  (time PTB200_LIB=build/variants/libptb200_qnodes.so python -m pytest tests -m gpu -q --durations=5) > gpurun_out/${tag}_qnodes.log 2>&1; tail -6 gpurun_out/${tag}_qnodes.log
fi
