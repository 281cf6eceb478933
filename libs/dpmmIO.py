"""Input and output of a BnpC run -- the callers on either side of the MCMC hot path (reference
cbg-ethz/BnpC v0.2.1 `libs/dpmmIO.py`).  Same function names and file
formats as the reference for everything `run_BnpC.py` needs without plotting (`-np`): the matrix
loader, the termination rule, the estimator driver and the text writers (args.txt, errors.txt,
assignment.txt, genotypes_*.tsv, ARI / V-measure / Hamming files).  Plots (matplotlib / seaborn /
graphviz; `libs/plotting.py`) are out of scope: `run_BnpC.py` says so and goes on.
"""
import os
from datetime import timedelta

import numpy as np
import pandas as pd

try:
    import libs.utils as ut
except ImportError:                                          # run from inside libs/, as the reference allows
    import utils as ut


# ------------------------------------------------------------------------------------ input
def _is_matrix_value(token):
    try:
        return float(token) in (0, 1, 2, 3)
    except ValueError:
        return token == ' '


CACHE_SUFFIX = '.bnpc_planes.npz'


def _cache_path(in_file):
    return in_file + CACHE_SUFFIX


def _load_cache(in_file):
    """The packed copy of a matrix file written by an earlier load (two bit-planes, the wire format
    of the kernels: plane1 bit = entry is 1, plane0 bit = entry is 0, neither = missing) -- valid
    while the text file has the size and modification time recorded in it."""
    path = _cache_path(in_file)
    if os.environ.get('BNPC_NO_CACHE') or not os.path.exists(path):
        return None
    try:
        st = os.stat(in_file)
        with np.load(path, allow_pickle=False) as z:
            if int(z['src_size']) != st.st_size or int(z['src_mtime_ns']) != st.st_mtime_ns:
                return None
            rows, cols = (int(v) for v in z['shape'])
            one = np.unpackbits(z['plane1'], axis=1, count=cols).astype(bool)
            zero = np.unpackbits(z['plane0'], axis=1, count=cols).astype(bool)
            values = np.full((rows, cols), np.nan)
            values[one] = 1.0
            values[zero] = 0.0
            return values, (z['row_names'], z['col_names'])
    except (OSError, KeyError, ValueError):
        return None


def _storable(names):
    a = np.asarray(names)
    return a if a.dtype.kind in 'iuf' else a.astype(str)


def _save_cache(in_file, values, names):
    """values: the parsed file matrix (rows x columns as in the file, {0, 1, NaN})"""
    if os.environ.get('BNPC_NO_CACHE'):
        return
    try:
        # a directory marked read-only stays untouched (also for root, whom the kernel would let write)
        if not os.stat(os.path.dirname(os.path.abspath(in_file))).st_mode & 0o200:
            return
        st = os.stat(in_file)
        tmp = _cache_path(in_file) + f'.tmp{os.getpid()}'
        with open(tmp, 'wb') as f:
            np.savez(f, plane1=np.packbits(values == 1, axis=1), plane0=np.packbits(values == 0, axis=1),
                     shape=np.array(values.shape), src_size=st.st_size, src_mtime_ns=st.st_mtime_ns,
                     row_names=_storable(names[0]), col_names=_storable(names[1]))
        os.replace(tmp, _cache_path(in_file))
    except OSError:                                          # read-only input directory: no cache
        pass


def _parse_matrix_file(in_file):
    """the text parse of libs/dpmmIO.py:27-86 -> (values rows x columns as in the file, names)"""
    with open(in_file, 'r') as f:
        head = [f.readline().strip() for _ in range(5)]
    head = [h for h in head if h]
    tabs, commas, blanks = head[0].count('\t'), head[0].count(','), head[0].count(' ')
    sep = '\t' if tabs > blanks and tabs > commas else (',' if commas > blanks else ' ')
    header_row = not all(_is_matrix_value(t) for t in head[0].split(sep))
    body = head[1:] if header_row else head
    index_col = not all(_is_matrix_value(line.split(sep)[0]) for line in body)
    df = pd.read_csv(in_file, sep=sep, index_col=0 if index_col else None, header=0 if header_row else None,
                     na_values=[3, ' '] if (index_col and header_row) else None)
    df = df.astype(float)
    values = df.values.copy()
    values[values == 3] = np.nan
    values[values == 2] = 1
    return values, (df.index.values, df.columns.values)


def load_data(in_file, transpose=True, get_names=False):
    """libs/dpmmIO.py:27-98.  Text matrix of 0 | 1 | 2 (homozygous, read as 1) | 3 or empty (missing),
    separated by tabs, commas or blanks, with an optional header row and an optional index column.
    Returns float64 [cells, mutations] with NaN for missing (after the default transpose: files are
    mutations x cells), i.e. what the model constructors take; they pack it into bit-planes on the
    device (`bnpc_pack_planes`).  The parse of a 100k x 1k text file dominates the start-up of a
    run: the parsed matrix is kept next to the input as two packed bit-planes (`<file>.bnpc_planes.npz`,
    1/32 of the float64 size) and read back from there while the text file is unchanged
    (BNPC_NO_CACHE=1 switches that off)."""
    cached = _load_cache(in_file)
    if cached is None:
        values, names = _parse_matrix_file(in_file)
        _save_cache(in_file, values, names)
    else:
        values, names = cached
    row_names, col_names = names
    if transpose:
        values, row_names, col_names = values.T.copy(), col_names, row_names
    if get_names:
        return values, (np.asarray(row_names), np.asarray(col_names))
    return values


def load_txt(path):
    """libs/dpmmIO.py:101-112: an assignment vector, either a one-line list of integers or the
    `Assignment` column of an assignment.txt written by save_assignments."""
    try:
        df = pd.read_csv(path, sep='\t', index_col=False)
        tokens = df.at[0, 'Assignment'].split(' ')
    except (ValueError, KeyError):
        with open(path, 'r') as f:
            tokens = f.read().split()
    return [int(t) for t in tokens]


def process_sim_folder(args, suffix=''):
    """libs/dpmmIO.py:119-154 for the file case: pick up data_raw.csv next to the input."""
    if os.path.isdir(args.input):
        in_dir = args.input
        args.input = os.path.join(in_dir, f'data{suffix}.csv')
        if getattr(args, 'transpose', False):
            args.true_clusters = os.path.join(in_dir, 'attachments.txt')
    else:
        in_dir = os.path.dirname(args.input)
    raw = os.path.join(in_dir, 'data_raw.csv')
    if os.path.exists(raw):
        args.true_data = raw


def _get_mcmc_termination(args):
    """libs/dpmmIO.py:157-169."""
    if args.runtime > 0:
        span = timedelta(minutes=args.runtime)
        return (args.time[0] + span, args.time[0] + args.burn_in * span), f'for {args.runtime} mins'
    if args.lugsail > 0:
        return (args.lugsail, 0), f'until PSRF < {args.lugsail:.4f}'
    return (args.steps, int(args.steps * args.burn_in)), f'for {args.steps} steps'


def _get_out_dir(args, prefix=''):
    """libs/dpmmIO.py:172-192."""
    if args.output:
        is_file = any(args.output.endswith(e) for e in ('.txt', '.gv', '.csv'))
        out_dir = os.path.dirname(args.output) if is_file else args.output
    else:
        base = os.path.join(os.path.dirname(args.input), f'BnpC_{args.time[0]:%Y%m%d_%H:%M:%S}{prefix}')
        out_dir, i = base, 1
        while os.path.exists(out_dir):
            out_dir = f'{base}_{i}'
            i += 1
    os.makedirs(out_dir, exist_ok=True)
    return out_dir


# ----------------------------------------------------------------------------- estimators
def _infer_results(args, results, data):
    """libs/dpmmIO.py:199-225: PSRF of the ML traces, then every requested estimator (posterior:
    GPU co-clustering + MPEAR, libs/utils.py; ML / MAP: the best sample)."""
    args.PSRF = ut.get_lugsail_batch_means_est([(r['ML'], r['burn_in']) for r in results])
    args.steps = [r['ML'].size for r in results]
    if isinstance(args.estimator, str):
        args.estimator = [args.estimator]
    inferred = {i: {} for i in range(args.chains)} if args.single_chains else {0: {}}
    for est in args.estimator:
        if est == 'posterior':
            found = ut.get_latents_posterior(results, data, args.single_chains)
        else:
            found = ut.get_latents_point(results, est, data, args.single_chains)
        for i, one in enumerate(found):
            inferred[i][est] = one
    if not args.single_chains:
        inferred['mean'] = inferred.pop(0)
    return inferred


# --------------------------------------------------------------------------------- stdout
def show_MCMC_summary(args, results):
    total = args.time[1] - args.time[0]
    steps = int(np.sum([r['ML'].size for r in results]))
    print(f'\nClustering time:\t{total}\t({steps} steps over {len(results)} chains, '
          f'{steps / max(total.total_seconds(), 1e-9):.1f} steps/s)')
    if len(results) > 1:
        print(f'Lugsail PSRF:\t\t{args.PSRF:.5f}\n')


def show_assignments(data, names=np.array([])):
    for chain, per_est in data.items():
        for est, found in per_est.items():
            ids, sizes = np.unique(found['assignment'], return_counts=True)
            print(f'Chain {chain} - {est} clusters\t#{ids.size}: sizes '
                  + ' '.join(str(s) for s in sorted(sizes, reverse=True)))


def show_latents(data):
    for chain, per_est in data.items():
        for est, found in per_est.items():
            if est == 'posterior':
                print(f'Chain {chain} - {est}:\tFN {found["FN"][0]:.4f}+-{found["FN"][1]:.4f} '
                      f'(data {found["FN_geno"]:.4f})\tFP {found["FP"][0]:.6f}+-{found["FP"][1]:.6f} '
                      f'(data {found["FP_geno"]:.6f})\talpha {found["a"][0]:.1f}+-{found["a"][1]:.1f}')
            else:
                print(f'Chain {chain} - {est}:\tstep {found["step"]}\tFN {found["FN"]:.4f} '
                      f'(data {found["FN_geno"]:.4f})\tFP {found["FP"]:.6f} (data {found["FP_geno"]:.6f})\t'
                      f'alpha {found["a"]:.1f}')


# -------------------------------------------------------------------------------- writers
def save_run(inferred, args, out_dir, names):
    save_config(args, out_dir)
    save_errors(inferred, args, out_dir)
    save_assignments(inferred, args, out_dir)
    save_geno(inferred, out_dir, names[1])


def save_config(args, out_dir, out_file='args.txt'):
    """libs/dpmmIO.py:429-451: one `key: value` line per argument; fixed error rates hide the
    prior settings and vice versa."""
    cfg = dict(args) if isinstance(args, dict) else dict(vars(args))
    cfg['time'] = [f'{t:%Y%m%d_%H:%M:%S}' for t in cfg['time']]
    for rate in ('falseNegative', 'falsePositive'):
        if cfg[rate] > 0:
            cfg.pop(f'{rate}_mean', None)
            cfg.pop(f'{rate}_std', None)
        else:
            cfg.pop(rate, None)
    with open(os.path.join(out_dir, out_file), 'w') as f:
        for key, val in cfg.items():
            f.write(f'{key}: {val}\n')


def _rows(data):
    for chain, per_est in data.items():
        for est, found in per_est.items():
            yield chain, est, found


def save_errors(data, args, out_dir):
    rows = []
    for chain, est, found in _rows(data):
        if est == 'posterior':
            rows.append([chain, est, f'{found["FN"][0]:.4f}+-{found["FN"][1]:.4f}', np.round(found['FN_geno'], 4),
                         f'{found["FP"][0]:.8f}+-{found["FP"][1]:.8f}', np.round(found['FP_geno'], 8)])
        else:
            rows.append([chain, est, np.round(found['FN'], 4), np.round(found['FN_geno'], 4),
                         np.round(found['FP'], 8), np.round(found['FP_geno'], 8)])
    df = pd.DataFrame(rows, columns=['chain', 'estimator', 'FN_model', 'FN_data', 'FP_model', 'FP_data'])
    df.to_csv(os.path.join(out_dir, 'errors.txt'), index=False, sep='\t')


def save_assignments(data, args, out_dir):
    rows = [[chain, est, ' '.join(str(v) for v in found['assignment'])] for chain, est, found in _rows(data)]
    pd.DataFrame(rows, columns=['chain', 'estimator', 'Assignment']).to_csv(
        os.path.join(out_dir, 'assignment.txt'), index=False, sep='\t')


def save_geno(data, out_dir, names=np.array([])):
    """libs/dpmmIO.py:491-511: genotypes_<est>_<chain>.tsv (rounded, int) and, for continuous
    genotypes, genotypes_cont_<est>_<chain>.tsv (4 decimals)."""
    for chain, est, found in _rows(data):
        geno = found['genotypes'].copy()
        if names.size == geno.index.size:
            geno.index = names
        if not (geno.round() == geno).all().all():
            geno.round(4).to_csv(os.path.join(out_dir, f'genotypes_cont_{est}_{chain:0>2}.tsv'), sep='\t')
        geno.round().astype(int).to_csv(os.path.join(out_dir, f'genotypes_{est}_{chain:0>2}.tsv'), sep='\t')


def _metric_frame(data, true_cl, name, score):
    rows = [[chain, est, score(found['assignment'], true_cl)] for chain, est, found in _rows(data)]
    return pd.DataFrame(rows, columns=['chain', 'estimator', name])


def save_v_measure(data, true_cl, out_dir):
    _metric_frame(data, true_cl, 'V-measure', ut.get_v_measure).to_csv(
        os.path.join(out_dir, 'V_measure.txt'), index=False, sep='\t')


def save_ARI(data, true_cl, out_dir):
    _metric_frame(data, true_cl, 'ARI', ut.get_ARI).to_csv(os.path.join(out_dir, 'ARI.txt'), index=False, sep='\t')


def save_hamming_dist(data, true_data, out_dir):
    rows = [[chain, est, 1 - ut.get_hamming_dist(found['genotypes'], true_data) / true_data.size]
            for chain, est, found in _rows(data)]
    pd.DataFrame(rows, columns=['chain', 'estimator', '1 - norm Hamming distance']).to_csv(
        os.path.join(out_dir, 'hamming_distance.txt'), index=False, sep='\t')
