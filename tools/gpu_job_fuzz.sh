#!/bin/bash
# parity fuzz on the trimmed kernels (new seeds): Binet strict / hybrid vs oracle, full frames, Kerr
mkdir -p gpurun_out
timeout 900 python tools/parity_fuzz.py 31 32 > gpurun_out/r2_parity_fuzz_binet_seed31.log 2>&1; echo "binet fuzz rc=$?"; tail -4 gpurun_out/r2_parity_fuzz_binet_seed31.log
timeout 900 python tools/parity_fuzz_frames.py 32 > gpurun_out/r2_parity_fuzz_frames_seed32.log 2>&1; echo "frames fuzz rc=$?"; tail -3 gpurun_out/r2_parity_fuzz_frames_seed32.log
timeout 900 python tools/parity_fuzz_kerr.py 33 > gpurun_out/r2_parity_fuzz_kerr_seed33.log 2>&1; echo "kerr fuzz rc=$?"; tail -4 gpurun_out/r2_parity_fuzz_kerr_seed33.log
