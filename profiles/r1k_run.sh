#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python profiles/sweep.py "LIB=libSpirit.so" "LIB=libSpirit_pf3.so" "LIB=libSpirit_pf5.so" "LIB=libSpirit_pf8.so" "LIB=libSpirit.so" "LIB=libSpirit_pf3.so" "LIB=libSpirit_pf5.so" > gpurun_out/r1k_sweep.txt 2>&1
cat gpurun_out/r1k_sweep.txt
