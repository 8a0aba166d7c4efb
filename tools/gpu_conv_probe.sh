#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/conv_probe.py backbone.s3.b5.c1 backbone.s3.b5.c3 backbone.s4.b5.c1 backbone.s4.b5.c2 backbone.s4.b5.c3 backbone.s4.b5.se.fc2 seg.d2 seg.d7.p00 > gpurun_out/conv_probe.log 2>&1
tail -12 gpurun_out/conv_probe.log
