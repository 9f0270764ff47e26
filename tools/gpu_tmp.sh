# scratch script for one-off gpurun calls (overwritten freely); the maintained entry point is tools/gpu_full.sh
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
