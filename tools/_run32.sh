export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
python bench.py 2> gpurun_out/bench.err | tail -1 | tee gpurun_out/bench.json | cut -c1-3500
tail -2 gpurun_out/bench.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
