"""Runner helpers -- mirror of the hot-path parts of src/utils/main_util.py (:14-26 JSON override,
:43-62 distributed init, warm-up LR schedule used by mimic_runner.py:43-46)."""
import json
import os

import torch
import torch.distributed as dist


def overwrite_dict(org_dict, sub_dict):
    for sub_key, sub_value in sub_dict.items():
        if sub_key in org_dict and isinstance(sub_value, dict):
            overwrite_dict(org_dict[sub_key], sub_value)
        else:
            org_dict[sub_key] = sub_value


def overwrite_config(config, json_str):
    overwrite_dict(config, json.loads(json_str))


def init_distributed_mode(world_size=1, dist_url='env://'):
    """One process per GPU, NCCL over NVLink (main_util.py:43-62).  Returns (distributed, device_ids)."""
    if 'RANK' in os.environ and 'WORLD_SIZE' in os.environ:
        rank = int(os.environ['RANK'])
        world_size = int(os.environ['WORLD_SIZE'])
        device_id = int(os.environ.get('LOCAL_RANK', 0))
    elif 'SLURM_PROCID' in os.environ:
        rank = int(os.environ['SLURM_PROCID'])
        device_id = rank % max(torch.cuda.device_count(), 1)
    else:
        print('Not using distributed mode')
        return False, None
    if world_size <= 1:
        return False, None
    backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    if torch.cuda.is_available():
        torch.cuda.set_device(device_id)
    dist.init_process_group(backend=backend, init_method=dist_url, world_size=world_size, rank=rank)
    dist.barrier()
    if rank != 0:  # main_util.py:29-40: print only on the master
        import builtins
        builtin_print = builtins.print

        def quiet(*args, **kwargs):
            if kwargs.pop('force', False):
                builtin_print(*args, **kwargs)
        builtins.print = quiet
    return True, [device_id]


def warmup_lr_scheduler(optimizer, warmup_iters, warmup_factor):
    def f(x):
        if x >= warmup_iters:
            return 1
        alpha = float(x) / warmup_iters
        return warmup_factor * (1 - alpha) + alpha
    return torch.optim.lr_scheduler.LambdaLR(optimizer, f)
