# call AA (2 GPUs): BASELINE config #5's delivery on hardware - 4K frames decoded on both ranks, packed to u8 on the device,
# gathered to rank 0 over NCCL - and the plain 2-GPU bench line
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02aa_topo.txt 2>&1
NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --workload synth4k --frames-per-step 8 --gather u8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02aa_gather.json 2> gpurun_out/r02aa_gather.err
grep -m3 "NVLS\|via P2P\|NET/\|Channel 00" gpurun_out/r02aa_gather.err | cut -c1-200
tail -1 gpurun_out/r02aa_gather.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('gather line: value', round(d['value']), 'e2e', round(d['e2e']['value']), 'n', d['n_gpus']); print(json.dumps(d['gather'], indent=1))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02aa_bench2.json 2> gpurun_out/r02aa_bench2.err
tail -1 gpurun_out/r02aa_bench2.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('2-GPU synth8k: value', round(d['value']), 'e2e', round(d['e2e']['value']), 'u8', round(d['e2e_u8']['value']))"
tail -2 gpurun_out/r02aa_bench2.err
