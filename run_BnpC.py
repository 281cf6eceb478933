#!/usr/bin/env python3
"""Command line of BnpC on the B200 path: the options of the reference's `run_BnpC.py`
(cbg-ethz/BnpC v0.2.1, run_BnpC.py:13-196) with the same names, defaults and meaning; chains run as
host threads + CUDA streams (or one rank per GPU under torchrun) instead of forked processes.
Plots are not produced (no matplotlib in this build): the run behaves as with `-np`.

    python run_BnpC.py example.csv -n 8 -s 5000 -e posterior MAP -o out/
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 run_BnpC.py data.csv -n 8
"""
import argparse
from datetime import datetime

import libs.dpmmIO as io
from libs.MCMC import MCMC, dist_info


def _in_range(lo, hi, open_ends):
    def check(val):
        val = float(val)
        bad = (val <= lo or val >= hi) if open_ends else (val < lo or val > hi)
        if bad:
            brackets = '0 < x < 1' if open_ends else f'{lo} <= x <= {hi}'
            raise argparse.ArgumentTypeError(f'Invalid value: {val}. Values need to be {brackets}')
        return val
    return check


def parse_args(argv=None):
    ratio, percent, psrf = _in_range(0, 1, True), _in_range(0, 1, False), _in_range(1, 1.5, False)
    p = argparse.ArgumentParser(prog='BnpC', usage='python3 run_BnpC.py <DATA> [options]',
                                description='*** Clustering of single cell data based on a Dirichlet process. ***')
    p.add_argument('--version', action='version', version='0.2.1 (bnpc-b200)')
    p.add_argument('input', help='Path to the input matrix (n cells x m mutations after the default transpose; '
                                 '1|0 entries, missing values as 3 or empty; blank-, tab- or comma-separated).')
    p.add_argument('-t', '--transpose', action='store_false', help='Transpose the input matrix. Default = True.')
    p.add_argument('--debug', action='store_true', default=False, help='Run a single chain in the main thread.')

    m = p.add_argument_group('model')
    m.add_argument('-FN', '--falseNegative', type=float, default=-1,
                   help='Fixed false negative rate; if > 0 (together with -FP) error rates are not learned.')
    m.add_argument('-FP', '--falsePositive', type=float, default=-1, help='Fixed false positive rate.')
    m.add_argument('-FN_m', '--falseNegative_mean', type=ratio, default=0.2, help='Prior mean of FN. Default = 0.2.')
    m.add_argument('-FN_sd', '--falseNegative_std', type=ratio, default=0.1, help='Prior sd of FN. Default = 0.1.')
    m.add_argument('-FP_m', '--falsePositive_mean', type=ratio, default=0.01, help='Prior mean of FP. Default = 0.01.')
    m.add_argument('-FP_sd', '--falsePositive_std', type=ratio, default=0.01, help='Prior sd of FP. Default = 0.01.')
    m.add_argument('-ap', '--DPa_prior', type=float, nargs=2, default=[-1, -1],
                   help='Gamma(a, b) prior of the concentration parameter; negative = (sqrt(cells), 1).')
    m.add_argument('-pp', '--param_prior', type=float, nargs=2, default=[.25, .25],
                   help='Beta(a, b) prior of the cluster parameters. Default = [.25, .25].')
    m.add_argument('-fa', '--fixed_assignment', type=str, default='',
                   help='File with a fixed assignment (only parameters are sampled).')

    c = p.add_argument_group('MCMC')
    c.add_argument('-n', '--chains', type=int, default=1, help='Number of chains. Default = 1.')
    c.add_argument('-s', '--steps', type=int, default=5000, help='Steps per chain. Default = 5000.')
    c.add_argument('-r', '--runtime', type=int, default=-1, help='Runtime in minutes (overrides -s).')
    c.add_argument('-ls', '--lugsail', type=psrf, default=-1,
                   help='Run until the lugsail PSRF of the chains is below this cutoff (1 <= x <= 1.5).')
    c.add_argument('-b', '--burn_in', type=percent, default=0.33, help='Burn-in fraction. Default = 0.33.')
    c.add_argument('-cup', '--conc_update_prob', type=percent, default=0.25,
                   help='Probability of updating the concentration parameter per step. Default = 0.25.')
    c.add_argument('-eup', '--error_update_prob', type=percent, default=0.25,
                   help='Probability of updating the error rates per step. Default = 0.25.')
    c.add_argument('-smp', '--split_merge_prob', type=percent, default=0.33,
                   help='Probability of a split/merge move per step. Default = 0.33.')
    c.add_argument('-sms', '--split_merge_steps', type=int, default=3,
                   help='Restricted Gibbs scans per split/merge move. Default = 3.')
    c.add_argument('-smr', '--split_merge_ratios', type=percent, nargs=2, default=[0.75, 0.25],
                   help='Ratio of splits/merges. Default = 0.75:0.25')
    c.add_argument('-e', '--estimator', type=str, default='posterior', nargs='+', choices=['posterior', 'ML', 'MAP'],
                   help='Estimator(s) for the inferred latent variables. Default = posterior.')
    c.add_argument('-sc', '--single_chains', action='store_true', default=False,
                   help='Infer the latent variables per chain instead of over all chains.')
    c.add_argument('--seed', type=int, default=-1, help='Seed of the chain seeds. Default = random.')

    o = p.add_argument_group('output')
    o.add_argument('-o', '--output', type=str, default='', help='Output directory. Default = next to the input.')
    o.add_argument('-v', '--verbosity', type=int, default=1, choices=[0, 1, 2], help='Stdout verbosity. Default = 1.')
    o.add_argument('-np', '--no_plots', action='store_true', default=False, help='Accepted; plots are never drawn.')
    o.add_argument('-tr', '--tree', type=str, default='', help='Accepted for compatibility (tree plots are not drawn).')
    o.add_argument('-tc', '--true_clusters', type=str, default='', help='True assignment: ARI and V-measure are written.')
    o.add_argument('-td', '--true_data', type=str, default='', help='True genotypes: the Hamming distance is written.')
    return p.parse_args(argv)


def generate_output(args, results, data_raw, names):
    """run_BnpC.py:203-241 without the plots."""
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
    if not args.no_plots and args.verbosity > 0:
        print('(plots are not part of this build: trace, genotype and similarity plots were skipped)')
    return out_dir


def main(args):
    """run_BnpC.py:244-292."""
    io.process_sim_folder(args, suffix='')
    data, names = io.load_data(args.input, transpose=args.transpose, get_names=True)
    assert data.size > 0, f'Could not read data from file: {args.input}'
    if args.falsePositive > 0 and args.falseNegative > 0:
        args.error_update_prob = 0
        import libs.CRP as CRP
        model = CRP.CRP(data, DP_alpha=args.DPa_prior, param_beta=args.param_prior,
                        FN_error=args.falseNegative, FP_error=args.falsePositive)
    else:
        import libs.CRP_learning_errors as CRP
        model = CRP.CRP_errors_learning(data, DP_alpha=args.DPa_prior, param_beta=args.param_prior,
                                        FP_mean=args.falsePositive_mean, FP_sd=args.falsePositive_std,
                                        FN_mean=args.falseNegative_mean, FN_sd=args.falseNegative_std)
    args.time = [datetime.now()]
    run_var, run_str = io._get_mcmc_termination(args)
    mcmc = MCMC(model, sm_prob=args.split_merge_prob, dpa_prob=args.conc_update_prob,
                error_prob=args.error_update_prob, sm_ratios=args.split_merge_ratios,
                sm_steps=args.split_merge_steps)
    rank = dist_info()[0]
    if args.verbosity > 0 and rank == 0:
        print(model)
        print(mcmc)
        print(f'Run MCMC with ({args.chains} chains {run_str}):')
    if args.debug:
        args.chains = 1
    mcmc.run(run_var, args.seed, args.chains, args.verbosity, args.fixed_assignment, args.debug)
    if rank != 0:
        return None                                   # under torchrun rank 0 holds the gathered traces
    args.chain_seeds = mcmc.get_seeds()
    results = mcmc.get_results()
    args.time.append(datetime.now())
    return generate_output(args, results, data, names)


if __name__ == '__main__':
    main(parse_args())
