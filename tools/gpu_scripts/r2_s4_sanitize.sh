#!/bin/bash
# fourth session: compute-sanitizer over the kernels this session added / changed
cd "$GRAFT_REPO_ROOT"
K="groupnorm_statistics or attention_on_unfolded or attention_with_output_mapping or inference_shortcut or tc_mlp_chain or test_attention"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -x -q -m gpu -k "$K" > gpurun_out/r02s4_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02s4_memcheck.log; grep -c "Invalid\|misaligned" gpurun_out/r02s4_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests -x -q -m gpu -k "groupnorm_statistics or attention_on_unfolded or attention_with_output_mapping" > gpurun_out/r02s4_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r02s4_racecheck.log
