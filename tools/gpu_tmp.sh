timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "well_formed or utf16be_twins" 2>&1 | tail -4
timeout 200 simdutf_b200/build/with_b200/to_well_formed_utf16_tests -a b200 > gpurun_out/ref_wf.log 2>&1; echo "to_well_formed_utf16_tests rc=$? OK=$(grep -c ' OK' gpurun_out/ref_wf.log)"
bash tools/gpu_ncu_one.sh utf32to8 k_elem_transcode r01_elem_u32to8
bash tools/gpu_ncu_one.sh u8tol1 k_elem_transcode r01_elem_u8tol1
