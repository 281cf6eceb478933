#!/usr/bin/env python3
"""Command line of BnpC on the B200 path.

The option names, defaults and meaning are those of the reference's command line (cbg-ethz/BnpC
v0.2.1, run_BnpC.py:13-196), declared here as one table; the flow of `main` is the reference's
(run_BnpC.py:244-292): load the matrix, build the model, run the chains, infer the requested
estimators, write the text outputs.  Chains run as host threads + CUDA streams of one process, or
one rank per GPU under torchrun, instead of forked processes.  Plots are not part of this build:
a run behaves as with `-np`.

    python run_BnpC.py example.csv -n 8 -s 5000 -e posterior MAP -o out/
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 run_BnpC.py data.csv -n 8
"""
import argparse
from datetime import datetime

import libs.dpmmIO as io
from libs.MCMC import MCMC, dist_info


def _bounded(lo, hi, strict):
    """argparse type: a float inside (lo, hi) if strict, else inside [lo, hi]"""
    def convert(text):
        x = float(text)
        outside = (x <= lo or x >= hi) if strict else (x < lo or x > hi)
        if outside:
            span = f'{lo} < x < {hi}' if strict else f'{lo} <= x <= {hi}'
            raise argparse.ArgumentTypeError(f'Invalid value: {x}. Values need to be {span}')
        return x
    return convert


RATIO = _bounded(0, 1, True)
PERCENT = _bounded(0, 1, False)
PSRF = _bounded(1, 1.5, False)

# (group, flags, argparse keywords); `None` = top level
OPTIONS = [
    (None, ('input',), dict(help='matrix file: mutations x cells (cells x mutations with -t), entries 0|1, '
                                 'missing as 3 or empty, separated by blanks, tabs or commas')),
    (None, ('-t', '--transpose'), dict(action='store_false', help='do NOT transpose the input (default: transpose)')),
    (None, ('--debug',), dict(action='store_true', default=False, help='one chain in the main thread')),
    ('model', ('-FN', '--falseNegative'), dict(type=float, default=-1, help='fixed FN rate (with -FP: rates are not learned)')),
    ('model', ('-FP', '--falsePositive'), dict(type=float, default=-1, help='fixed FP rate')),
    ('model', ('-FN_m', '--falseNegative_mean'), dict(type=RATIO, default=0.2, help='prior mean of FN [0.2]')),
    ('model', ('-FN_sd', '--falseNegative_std'), dict(type=RATIO, default=0.1, help='prior sd of FN [0.1]')),
    ('model', ('-FP_m', '--falsePositive_mean'), dict(type=RATIO, default=0.01, help='prior mean of FP [0.01]')),
    ('model', ('-FP_sd', '--falsePositive_std'), dict(type=RATIO, default=0.01, help='prior sd of FP [0.01]')),
    ('model', ('-ap', '--DPa_prior'), dict(type=float, nargs=2, default=[-1, -1],
                                           help='Gamma(a, b) prior of the concentration; negative: (sqrt(cells), 1)')),
    ('model', ('-pp', '--param_prior'), dict(type=float, nargs=2, default=[.25, .25],
                                             help='Beta(a, b) prior of the cluster parameters [.25 .25]')),
    ('model', ('-fa', '--fixed_assignment'), dict(type=str, default='', help='file with an assignment that is kept fixed')),
    ('MCMC', ('-n', '--chains'), dict(type=int, default=1, help='chains [1]')),
    ('MCMC', ('-s', '--steps'), dict(type=int, default=5000, help='steps per chain [5000]')),
    ('MCMC', ('-r', '--runtime'), dict(type=int, default=-1, help='minutes to run (overrides -s)')),
    ('MCMC', ('-ls', '--lugsail'), dict(type=PSRF, default=-1, help='run until the lugsail PSRF falls below this cutoff')),
    ('MCMC', ('-b', '--burn_in'), dict(type=PERCENT, default=0.33, help='burn-in fraction [0.33]')),
    ('MCMC', ('-cup', '--conc_update_prob'), dict(type=PERCENT, default=0.25, help='P(update concentration) per step [0.25]')),
    ('MCMC', ('-eup', '--error_update_prob'), dict(type=PERCENT, default=0.25, help='P(update error rates) per step [0.25]')),
    ('MCMC', ('-smp', '--split_merge_prob'), dict(type=PERCENT, default=0.33, help='P(split/merge move) per step [0.33]')),
    ('MCMC', ('-sms', '--split_merge_steps'), dict(type=int, default=3, help='restricted Gibbs scans per split/merge move [3]')),
    ('MCMC', ('-smr', '--split_merge_ratios'), dict(type=PERCENT, nargs=2, default=[0.75, 0.25], help='split : merge [0.75 0.25]')),
    ('MCMC', ('-e', '--estimator'), dict(type=str, default='posterior', nargs='+', choices=['posterior', 'ML', 'MAP'],
                                         help='estimator(s) of the latent variables [posterior]')),
    ('MCMC', ('-sc', '--single_chains'), dict(action='store_true', default=False, help='estimators per chain')),
    ('MCMC', ('--seed',), dict(type=int, default=-1, help='seed of the chain seeds [random]')),
    ('output', ('-o', '--output'), dict(type=str, default='', help='output directory [next to the input]')),
    ('output', ('-v', '--verbosity'), dict(type=int, default=1, choices=[0, 1, 2], help='stdout verbosity [1]')),
    ('output', ('-np', '--no_plots'), dict(action='store_true', default=False, help='accepted; plots are never drawn')),
    ('output', ('-tr', '--tree'), dict(type=str, default='', help='accepted for compatibility (no tree plots)')),
    ('output', ('-tc', '--true_clusters'), dict(type=str, default='', help='true assignment: ARI and V-measure are written')),
    ('output', ('-td', '--true_data'), dict(type=str, default='', help='true genotypes: the Hamming distance is written')),
]


def parse_args(argv=None):
    parser = argparse.ArgumentParser(prog='BnpC', usage='python3 run_BnpC.py <DATA> [options]',
                                     description='Dirichlet-process clustering of single-cell mutation calls (B200 build)')
    parser.add_argument('--version', action='version', version='0.2.1 (bnpc-b200)')
    groups = {None: parser}
    for group, flags, kw in OPTIONS:
        if group not in groups:
            groups[group] = parser.add_argument_group(group)
        groups[group].add_argument(*flags, **kw)
    return parser.parse_args(argv)


def build_model(args, data):
    """fixed error rates when both -FN and -FP are given, learned otherwise (run_BnpC.py:251-262)"""
    common = dict(DP_alpha=args.DPa_prior, param_beta=args.param_prior)
    if args.falsePositive > 0 and args.falseNegative > 0:
        args.error_update_prob = 0
        from libs.CRP import CRP
        return CRP(data, FN_error=args.falseNegative, FP_error=args.falsePositive, **common)
    from libs.CRP_learning_errors import CRP_errors_learning
    return CRP_errors_learning(data, FP_mean=args.falsePositive_mean, FP_sd=args.falsePositive_std,
                               FN_mean=args.falseNegative_mean, FN_sd=args.falseNegative_std, **common)


def generate_output(args, results, data_raw, names):
    """estimators and text outputs (run_BnpC.py:203-241 without the plots)"""
    out_dir = io._get_out_dir(args)
    inferred = io._infer_results(args, results, data_raw)
    if args.verbosity > 0:
        io.show_MCMC_summary(args, results)
        io.show_assignments(inferred, names[0])
        io.show_latents(inferred)
        print(f'\nWriting output to: {out_dir}\n')
    io.save_run(inferred, args, out_dir, names)
    if args.true_clusters:
        truth = io.load_txt(args.true_clusters)
        io.save_v_measure(inferred, truth, out_dir)
        io.save_ARI(inferred, truth, out_dir)
    if args.true_data:
        io.save_hamming_dist(inferred, io.load_data(args.true_data, transpose=args.transpose), out_dir)
    if args.verbosity > 0 and not args.no_plots:
        print('(plots are not part of this build: trace, genotype and similarity plots were skipped)')
    return out_dir


def main(args):
    io.process_sim_folder(args, suffix='')
    data, names = io.load_data(args.input, transpose=args.transpose, get_names=True)
    assert data.size > 0, f'Could not read data from file: {args.input}'
    model = build_model(args, data)
    args.time = [datetime.now()]
    run_var, run_str = io._get_mcmc_termination(args)
    mcmc = MCMC(model, sm_prob=args.split_merge_prob, dpa_prob=args.conc_update_prob,
                error_prob=args.error_update_prob, sm_ratios=args.split_merge_ratios,
                sm_steps=args.split_merge_steps)
    rank = dist_info()[0]
    if rank == 0 and args.verbosity > 0:
        print(model)
        print(mcmc)
        print(f'Run MCMC with ({args.chains} chains {run_str}):')
    if args.debug:
        args.chains = 1
    mcmc.run(run_var, args.seed, args.chains, args.verbosity, args.fixed_assignment, args.debug)
    if rank != 0:
        return None                     # under torchrun rank 0 holds the gathered traces
    args.chain_seeds = mcmc.get_seeds()
    results = mcmc.get_results()
    args.time.append(datetime.now())
    return generate_output(args, results, data, names)


if __name__ == '__main__':
    main(parse_args())
