# usage: gpurun --gpus 2 -- 'bash tools/grid_check.sh': the grid tests (incl. the 2-rank one) and both forms of both directions timed at N = 2
mkdir -p gpurun_out/r02x
timeout 300 python -m pytest tests/test_gpu_grid.py -q -m gpu 2>&1 | tail -4
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 tools/grid_run.py > gpurun_out/r02x/grid_n2.json 2> gpurun_out/r02x/grid_n2.err
python -c "import json; d=json.loads(open('gpurun_out/r02x/grid_n2.json').read().strip().splitlines()[-1]); print({k:v for k,v in d.items() if 'check' in k or 'mode' in k}); [print(k, {a:round(b,3) for a,b in d[k].items() if a.startswith('ms_')}) for k in ('encode','encode_staged','decode','decode_peer')]" || tail -5 gpurun_out/r02x/grid_n2.err
