#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/conv_probe.py backbone.s0.b0.c1 backbone.s1.b0.c1 backbone.s2.b1.c1 backbone.s2.b1.c2 backbone.s3.b5.c1 backbone.s3.b5.c2 backbone.s3.b5.c3 backbone.s3.b5.se.fc1 backbone.s3.b5.se.fc2 backbone.s4.b5.c1 backbone.s4.b5.c2 backbone.s4.b5.c3 backbone.s4.b5.se.fc1 backbone.s4.b5.se.fc2 det.reg.0.pw neck.c1.conv3_up.pw lane.hidden > gpurun_out/conv_probe.log 2>&1
tail -22 gpurun_out/conv_probe.log
