#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -x -q -m gpu -k "w_pairs or fused_front or pools_in_epilogue or groupnorm_in_epilogue or stride2 or retrieval_backbone" > gpurun_out/r02s3_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/r02s3_memcheck.log; grep -c "Invalid\|misaligned" gpurun_out/r02s3_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests -x -q -m gpu -k "fused_front or groupnorm_in_epilogue" > gpurun_out/r02s3_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/r02s3_racecheck.log
