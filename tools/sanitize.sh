#!/bin/bash
# compute-sanitizer memcheck over the tests that stress ragged tiles, unaligned views, batches and the band pipeline.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/sanitize
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py tests/test_gpu_aux.py tests/test_gpu_gainmap.py tests/test_gpu_scale.py -q -m gpu -x \
    -k "wide_tiles or unaligned or batch or mismatched or staging or alpha_seam or premultiply_in_place or fast_kernels_are_used or formats_depths or rgb565 or zero_weight or image_entry_point or nan_and or every_loop or another_size or sample_aligned or identity_matrix or random or full_size or committed or every_transfer" \
    > gpurun_out/sanitize/memcheck.log 2>&1
echo "memcheck exit code $?" >> gpurun_out/sanitize/memcheck.log
tail -15 gpurun_out/sanitize/memcheck.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
