#!/usr/bin/env python
"""Drop-in for the reference's protein driver (scripts/directed_evolution.py:34-167) on the B200 path.

Same flags, same files read (`<protein_weights>/<protein>/{wt.fasta,potts.pkl,onehot_cnn_seed=*.pt,results-…-linear.pkl}`),
same files written (`config.txt`, `population.npy`, `pred_fitness_scores.npy`, `oracle_fitness_scores.npy`,
`potts_scores.npy`, `energy_scores.npy`, `energy_history.npy`, `fitness_history.npy`; reference :91-101).
Only `--sampler PPDE` with `--unsupervised_expert potts` (or `--energy_function supervised`) is on the hot path;
MSA-Transformer scoring needs the un-vendored `esm_one_hot` package and is not offered.

Multi-GPU: `torchrun --nproc-per-node N scripts/directed_evolution.py …` (chains are sharded, results gathered).
"""
import argparse
import datetime
import json
import os
import random
import sys
from pathlib import Path

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

from ppde_b200 import dist as D                                                    # noqa: E402
from ppde_b200 import weights as W                                                 # noqa: E402
from ppde_b200.energy import ProteinProductOfExperts, ProteinSupervised            # noqa: E402
from ppde_b200.ridge import AugmentedLinearRegression                              # noqa: E402
from ppde_b200.sampler import PPDE_PAS                                             # noqa: E402


def build_parser():
    p = argparse.ArgumentParser()
    g = p.add_argument_group('general')                                            # reference :113-146
    g.add_argument('--protein_weights', type=str, default='weights')
    g.add_argument('--results_path', type=str, default='results/proteins')
    g.add_argument('--protein', type=str, default='PABP_YEAST_Fields2013')
    g.add_argument('--hub_dir', type=str, default='.')
    g.add_argument('--msa_path', type=str, default='data/proteins/PABP_YEAST.a2m')
    g.add_argument('--msa_size', type=int, default=500)
    g.add_argument('--seed', type=int, default=1234567)
    g.add_argument('--device', type=str, default='cuda')
    g.add_argument('--log_every', type=int, default=50)
    g.add_argument('--run_signature', type=str, default='')
    g.add_argument('--n_iters', type=int, default=10000)
    g.add_argument('--n_chains', type=int, default=128)
    g.add_argument('--energy_lamda', type=float, default=5)
    g.add_argument('--energy_function', type=str, default='product_of_experts')
    g.add_argument('--unsupervised_expert', type=str, default='potts')
    g.add_argument('--sampler', type=str, default='PPDE')
    g.add_argument('--nmut_threshold', type=int, default=0)
    g.add_argument('--disable_MSA_transformer_scoring', action='store_true')
    g.add_argument('--paper_results', action='store_true', default=False)
    p.add_argument_group('ppde').add_argument('--ppde_pas_length', type=int, default=2)
    return p


def main(args):
    np.random.seed(args.seed)                                                       # reference :38-40
    random.seed(args.seed)
    torch.manual_seed(args.seed)
    if args.sampler != 'PPDE':
        raise SystemExit("only --sampler PPDE is on the B200 hot path (SURVEY.md §8)")
    if 'RANK' in os.environ and not torch.distributed.is_initialized():
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        torch.distributed.init_process_group('nccl')
        args.device = f"cuda:{int(os.environ.get('LOCAL_RANK', '0'))}"
    rank, _ = D.world()

    tag = "{}_{}".format(args.sampler, args.seed) if args.run_signature == '' else \
        "{}_{}_{}".format(args.sampler, args.run_signature, args.seed)
    results_path = Path(args.results_path, args.protein,
                        tag + "_" + datetime.datetime.now().strftime("%Y-%m-%d_%H-%M-%S"))
    if rank == 0:
        results_path.mkdir(parents=True, exist_ok=True)

    dataset = os.path.join(args.protein_weights, args.protein)
    if args.energy_function == 'product_of_experts':
        energy_func = ProteinProductOfExperts(args)
    elif args.energy_function == 'supervised':
        energy_func = ProteinSupervised(args)
    else:
        raise SystemExit(f"unknown --energy_function {args.energy_function}")
    energy_func = energy_func.to(args.device)
    m = energy_func.model

    # the oracle model needs the Potts expert; with the supervised energy a Potts-enabled twin is loaded for scoring only
    if m.has_potts:
        om, reg = m, getattr(energy_func, "reg_coef", 1.0)
    else:
        twin = ProteinProductOfExperts(argparse.Namespace(**{**vars(args), "unsupervised_expert": "potts"}))
        om, reg = twin.model, twin.reg_coef
    oracle = AugmentedLinearRegression.from_dataset(om, dataset, reg_potts=reg)

    seqs, _ = W.read_fasta(os.path.join(dataset, 'wt.fasta'))
    wt = torch.from_numpy(W.seq_to_aa(seqs[0]).astype(np.int64))
    initial_population = torch.nn.functional.one_hot(wt, 20).float()[None].to(m.device).repeat(args.n_chains, 1, 1)
    if rank == 0:
        print(f'WT protein energy: {energy_func.get_energy(initial_population[:1])[0].mean():.3f}')

    sampler = PPDE_PAS(args)
    best_samples, best_energy, best_fitness, energy_history, fitness_history, random_traj = \
        sampler.run(initial_population, args.n_iters, energy_func, oracle.potts.index_list[0],
                    oracle.potts.index_list[-1], oracle if om is m else oracle.__call__, args.log_every)

    aa_best = om.onehot_to_aa(best_samples)
    best_oracle = oracle.score_states(aa_best).cpu().numpy()
    potts_score = om.energy(aa_best, want_grad=False)[3].cpu().numpy()               # proteins_potts_score (metrics.py:14-19)
    if rank == 0:
        q = [0.2, 0.4, 0.6, 0.8, 1.0]
        print(f'energy quantiles: {np.quantile(best_energy, q)}')
        print(f'fitness quantiles: {np.quantile(best_fitness, q)}')
        print(f'oracle quantiles: {np.quantile(best_oracle, q)}')
        print(f'potts quantiles: {np.quantile(potts_score, q)}')
        with open(results_path / 'config.txt', 'w') as f:
            json.dump(args.__dict__, f, indent=2)
        np.save(results_path / 'population.npy', best_samples.detach().cpu().numpy())
        np.save(results_path / 'pred_fitness_scores.npy', best_fitness)
        np.save(results_path / 'oracle_fitness_scores.npy', best_oracle)
        np.save(results_path / 'potts_scores.npy', potts_score)
        np.save(results_path / 'energy_scores.npy', best_energy)
        np.save(results_path / 'energy_history.npy', energy_history)
        np.save(results_path / 'fitness_history.npy', fitness_history)
        print('done')
    return results_path


if __name__ == '__main__':
    main(build_parser().parse_args())
