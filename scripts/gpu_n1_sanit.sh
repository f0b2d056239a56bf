#!/bin/bash
bash scripts/gpu_multi_r2.sh 1
bash scripts/gpu_sanitizer.sh
