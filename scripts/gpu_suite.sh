#!/bin/bash
# The -m gpu suite under both trees (run under gpurun): default environment, then PTB_BVH=wide / binary (whichever is not
# the default), logs under gpurun_out/.
tag=${1:-suite}
(time python -m pytest tests -m gpu -q --durations=5) > gpurun_out/${tag}_default.log 2>&1; tail -6 gpurun_out/${tag}_default.log
(time PTB_BVH=wide python -m pytest tests -m gpu -q --durations=5) > gpurun_out/${tag}_wide.log 2>&1; tail -6 gpurun_out/${tag}_wide.log
