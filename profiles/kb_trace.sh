#!/bin/bash
for k in 16 64 128 200; do GSB_TRACE=1 python profiles/k1_probe.py $k 296 2; done
