#!/usr/bin/env python
"""Launcher with the reference's `main.py` behaviour on the B200 path: read `model.yaml` (the run's own copy when the run
directory exists), build `Unet3D` / `GaussianDiffusion` / `Trainer` from it, train (or resume from `--load-step`), then
sample videos for the target stress-strain curves (ref main.py:9-117).

    python main.py --run-name my_run                                     # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 main.py --run-name my_run

What the reference hard-codes at the top of its `main()` (run name, checkpoint step, number of predictions, guidance scale,
wandb user) are command-line options here; their defaults are the reference's values.  Data layout, run-directory layout and
output files are the reference's: `data/<reference_frame>/{training,validation}/`, `runs/<run>/model/model.yaml`,
`runs/<run>/model/step_<n>/checkpoint.pt`, `runs/<run>/<mode>/step_<n>/{gifs/,geometries.csv}`.
"""
from __future__ import annotations

import argparse
import os
from pathlib import Path

import yaml

CONFIG_KEYS = ('batch_size', 'learning_rate', 'selected_channels', 'train_timesteps', 'sampling_timesteps', 'use_dynamic_thres',
               'reference_frame', 'padding_mode', 'unet_dim', 'unet_attn_dim_head', 'unet_attn_heads', 'unet_resnet_groups',
               'unet_cond_attention', 'unet_cond_to_time', 'unet_temporal_att_cond', 'unet_use_sparse_linear_attn', 'per_frame_cond',
               'unet_cond_att_GRU', 'unet_cond_attention_tokens')


def load_config(path) -> dict:
    config = yaml.safe_load(Path(path).read_text())
    missing = [k for k in CONFIG_KEYS if k not in config]
    if missing:
        raise KeyError(f'{path}: missing configuration keys {missing}')
    return config


def build_model(config: dict, image_size: int = 96, num_frames: int = 11):
    """`Unet3D` + `GaussianDiffusion` from a model.yaml dictionary, with the keyword wiring of ref main.py:62-91."""
    from denoising_diffusion_pytorch import GaussianDiffusion, Unet3D
    channels = len(config['selected_channels'])
    model = Unet3D(dim=config['unet_dim'], dim_mults=(1, 2, 4, 8), channels=channels, attn_heads=config['unet_attn_heads'],
                   attn_dim_head=config['unet_attn_dim_head'], init_dim=None, init_kernel_size=7,
                   use_sparse_linear_attn=config['unet_use_sparse_linear_attn'], resnet_groups=config['unet_resnet_groups'],
                   cond_bias=True, cond_attention=config['unet_cond_attention'], cond_attention_tokens=config['unet_cond_attention_tokens'],
                   cond_att_GRU=config['unet_cond_att_GRU'], use_temporal_attention_cond=config['unet_temporal_att_cond'],
                   cond_to_time=config['unet_cond_to_time'], per_frame_cond=config['per_frame_cond'], padding_mode=config['padding_mode'])
    diffusion = GaussianDiffusion(model, image_size=image_size, channels=channels, num_frames=num_frames,
                                  timesteps=config['train_timesteps'], loss_type='l1', use_dynamic_thres=config['use_dynamic_thres'],
                                  sampling_timesteps=config['sampling_timesteps'])
    return model, diffusion


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument('--run-name', default='pretrained')
    ap.add_argument('--load-step', type=int, default=None, help="checkpoint step to resume from (default: 200000 for 'pretrained', else none)")
    ap.add_argument('--num-preds', type=int, default=1, help='predictions per conditioning')
    ap.add_argument('--guidance-scale', type=float, default=5.)
    ap.add_argument('--train-steps', type=int, default=200000)
    ap.add_argument('--save-and-sample-every', type=int, default=10000)
    ap.add_argument('--root', default='./', help='directory holding data/, runs/ and model.yaml')
    ap.add_argument('--targets', default=None, help='target curves (default: <root>/data/target_responses.csv)')
    ap.add_argument('--wandb-username', default=None)
    ap.add_argument('--preload-data', action='store_true', help='decode the whole GIF dataset into host memory before training')
    ap.add_argument('--device-dataset', action='store_true',
                    help='decode the training GIFs on the GPU once and serve every batch from HBM (device_dataset.py); default: host Dataset + DataLoader')
    ap.add_argument('--synthetic-data', action='store_true', help='train on in-memory random clips when the data folders are absent')
    args = ap.parse_args(argv)

    from videometamaterials_b200 import Accelerator, DistributedDataParallelKwargs
    from denoising_diffusion_pytorch import Trainer
    load_step = args.load_step if args.load_step is not None else (200000 if args.run_name == 'pretrained' else None)
    accelerator = Accelerator(mixed_precision='fp16', kwargs_handlers=[DistributedDataParallelKwargs(find_unused_parameters=True)],
                              log_with='wandb' if args.wandb_username is not None else None)
    root = args.root if args.root.endswith('/') else args.root + '/'
    run_dir = root + 'runs/' + args.run_name + '/'
    if os.path.exists(run_dir):
        if load_step is None:
            accelerator.print('Directory already exists, please change run_name to train new model or provide load_model_step')
            return 1
        config = load_config(run_dir + 'model/model.yaml')
    else:
        config = load_config(root + 'model.yaml')
        accelerator.wait_for_everyone()
        if accelerator.is_main_process:
            os.makedirs(run_dir + 'training')
            os.makedirs(run_dir + 'model')
            with open(run_dir + 'model/model.yaml', 'w') as f:
                yaml.dump(config, f)
        accelerator.wait_for_everyone()
    _, diffusion = build_model(config)
    data_dir = root + 'data/' + config['reference_frame'] + '/'
    if not args.synthetic_data:
        for sub in ('training/', 'validation/'):
            if not os.path.isdir(data_dir + sub):
                raise FileNotFoundError(f'{data_dir + sub} not found (the dataset is an external download; --synthetic-data trains on random clips)')
    trainer = Trainer(diffusion, folder=data_dir + 'training/', validation_folder=data_dir + 'validation/', results_folder=run_dir,
                      selected_channels=config['selected_channels'], train_batch_size=config['batch_size'], test_batch_size=config['batch_size'],
                      train_lr=config['learning_rate'], save_and_sample_every=args.save_and_sample_every, train_num_steps=args.train_steps,
                      ema_decay=0.995, log=True, null_cond_prob=0.1, per_frame_cond=config['per_frame_cond'],
                      reference_frame=config['reference_frame'], run_name=args.run_name, accelerator=accelerator,
                      wandb_username=args.wandb_username, preload_data=args.preload_data, synthetic_data=args.synthetic_data,
                      device_dataset=args.device_dataset)
    trainer.train(load_model_step=load_step, num_samples=3, num_preds=args.num_preds)
    trainer.eval_target(args.targets or root + 'data/target_responses.csv', guidance_scale=args.guidance_scale, num_preds=args.num_preds)
    return 0


if __name__ == '__main__':
    raise SystemExit(main())
