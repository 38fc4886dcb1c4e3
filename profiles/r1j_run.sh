#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python profiles/sweep.py "LIB=libSpirit.so" "LIB=libSpirit.so SPIRIT_B200_SC6_LC1=22" "LIB=libSpirit.so SPIRIT_B200_SC6_LC1=26" "LIB=libSpirit.so SPIRIT_B200_SC6_LC1=16" "LIB=libSpirit.so SPIRIT_B200_SC6_LC1=11" "LIB=libSpirit.so SPIRIT_B200_SC6_LC2=18" "LIB=libSpirit.so SPIRIT_B200_SC6_LC2=16" "LIB=libSpirit.so SPIRIT_B200_SC6_LC1=22 SPIRIT_B200_SC6_LC2=18" "LIB=libSpirit.so" > gpurun_out/r1j_sweep.txt 2>&1
cat gpurun_out/r1j_sweep.txt
