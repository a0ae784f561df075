R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 70 $R --master-port 29521 tools/scale_probe.py --tag default > gpurun_out/r2_probe8_a.json 2> gpurun_out/r2_probe8_a.err
NCCL_MAX_CTAS=2 timeout 70 $R --master-port 29522 tools/scale_probe.py --tag maxctas2 > gpurun_out/r2_probe8_b.json 2> gpurun_out/r2_probe8_b.err
NCCL_MAX_CTAS=32 NCCL_PROTO=LL timeout 70 $R --master-port 29523 tools/scale_probe.py --tag protoLL_ctas32 > gpurun_out/r2_probe8_c.json 2> gpurun_out/r2_probe8_c.err
HL_PDL=0 timeout 70 $R --master-port 29524 tools/scale_probe.py --tag nopdl > gpurun_out/r2_probe8_d.json 2> gpurun_out/r2_probe8_d.err
cat gpurun_out/r2_probe8_*.json | grep '^{'
