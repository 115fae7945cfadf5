# The driver's scaling run, reproduced on one 8 x B200 box: bench.py at N = 1, 2, 4, 8 (weak scaling of the contract step,
# the strong-scaling block, the end-to-end leg).  Output: gpurun_out/scale_N.json
set -x
python bench.py --gpus 1 --no-extra --no-cpu-baseline > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --no-extra --no-cpu-baseline > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
done
tail -n 2 gpurun_out/scale_*.err
