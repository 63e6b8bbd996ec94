export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_.*_tc -s 4 -c 2 -f -o gpurun_out/attn_tc python tools/attn_prof.py > gpurun_out/ncu_attn.log 2>&1
tail -3 gpurun_out/ncu_attn.log
ls -la gpurun_out/*.ncu-rep
