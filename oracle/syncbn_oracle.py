"""TEST INFRASTRUCTURE ONLY (never imported by ams_b200): numpy restatement of the cross-rank BatchNorm arithmetic that
the BN finalize kernels implement for data parallel (ams_b200/csrc/bn_kernels.cu, `syncbn_exchange_block`,
`bn_finalize_kernel`, `bn_bwd_finalize_kernel`, `imgpool_bn_kernel`).

What it must equal is the reference's single-process semantics: FusedBatchNormV3(is_training=True) over the WHOLE batch
(checkpoints/*/model.meta; update ops run with the train step, utils/graph_utils.py:457) -- i.e. the same function as
oracle/student_oracle.py `batch_norm(..., mode='batch')` applied to the concatenation of the ranks' shards.  Parity
unpinned at the TF level like the rest of the oracle (TensorFlow 1.15 is not installable here); pinned against torch
autograd on the concatenated batch in tests/test_oracle.py.
"""
import numpy as np


def forward_sums(z):
    """per-rank pair the forward finalize pushes to its peers: (sum z, sum z^2) over the rank's rows, fp64, per channel"""
    z = np.asarray(z, np.float64)
    return z.sum(0), (z * z).sum(0)


def global_stats(pairs, rows_per_rank, eps):
    """mean, biased var, rstd from the ranks' pairs added in rank order with n = rows_per_rank * world (bn_finalize_kernel)"""
    s = np.zeros_like(pairs[0][0])
    q = np.zeros_like(pairs[0][1])
    for a, b in pairs:                       # rank order: every rank adds the same sequence -> bit-identical statistics
        s = s + a
        q = q + b
    n = float(rows_per_rank * len(pairs))
    mean = s / n
    var = np.maximum(q / n - mean * mean, 0.0)
    meanf, varf = mean.astype(np.float32), var.astype(np.float32)
    rstd = (1.0 / np.sqrt(varf + np.float32(eps))).astype(np.float32)
    unbiased = (var * (n / max(n - 1.0, 1.0))).astype(np.float32)
    return meanf, varf, rstd, unbiased


def backward_sums(g, z):
    """per-rank pair of the backward finalize: (sum g, sum g*z) with g already masked by the activation"""
    g, z = np.asarray(g, np.float64), np.asarray(z, np.float64)
    return g.sum(0), (g * z).sum(0)


def backward_coefficients(pairs, local_pair, rows_per_rank, mean, rstd, gamma):
    """dz = A*g + B*z + Cc from the GLOBAL sums; d_gamma / d_beta from the LOCAL ones (the gradient allreduce adds the
    ranks) -- bn_bwd_finalize_kernel"""
    mean, rstd, gamma = (np.asarray(x, np.float64) for x in (mean, rstd, gamma))
    s1 = np.zeros_like(pairs[0][0])
    sz = np.zeros_like(pairs[0][1])
    for a, b in pairs:
        s1 = s1 + a
        sz = sz + b
    n = float(rows_per_rank * len(pairs))
    s2 = rstd * (sz - mean * s1)
    A = gamma * rstd
    B = -gamma * rstd * rstd * s2 / n
    Cc = -gamma * rstd * (s1 / n - mean * rstd * s2 / n)
    l1, lz = local_pair
    d_gamma = (rstd * (lz - mean * l1)).astype(np.float32)
    d_beta = l1.astype(np.float32)
    return A.astype(np.float32), B.astype(np.float32), Cc.astype(np.float32), d_gamma, d_beta


def pooled_mean_var(local_pairs, n_per_rank):
    """image-pooling BN (statistics over the batch dimension only): every rank contributes (sum, centred sum of squares
    around ITS mean); pooled with the parallel-variance formula (imgpool_bn_kernel)"""
    world = len(local_pairs)
    cnt = float(n_per_rank * world)
    S = np.zeros_like(local_pairs[0][0])
    for s, _ in local_pairs:
        S = S + s
    mean = S / cnt
    M2 = np.zeros_like(S)
    for s, m2 in local_pairs:
        d = s / n_per_rank - mean
        M2 = M2 + (m2 + n_per_rank * d * d)
    return mean, M2 / cnt
