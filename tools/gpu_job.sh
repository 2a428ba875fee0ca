#!/bin/bash
# tools/gpu_job.sh TAG 'cmd; cmd; ...' -- run on the GPU box (through gpurun), everything into
# gpurun_out/<TAG>.log so that a call cut off by the shell tool can still be read.
TAG=$1; shift
mkdir -p gpurun_out
( eval "$@" ) > gpurun_out/$TAG.log 2>&1
tail -n 60 gpurun_out/$TAG.log
