"""Packed estimator buffer and walker sharding -- the host-side logic around the one collective.

The reference reduces seven arrays to the root rank, one MPI_Reduce each, and divides by the number
of ranks (src/TDVMC.cpp:1182-1188, src/MPIMethods.h:132-206, 329-362).  Here every rank holds SUMS
over its local samples in one packed buffer

    [ S (P*P) | F_R (P) | F_I (P) | O (P) | E_R | E_I | other (n_other) | n_acc | n_trials | n_samples ]

which is all-reduced once (NCCL inside libtdvmc_b200.so; ``torch.distributed`` here for the
host-side mirror used in the CPU tests) and divided by the global sample count.  Because every rank
draws the same number of samples per walker this equals the reference's mean over ranks.
"""
from dataclasses import dataclass

import numpy as np


def shard_walkers(n_total, rank, world):
    """Contiguous, balanced split of global walker ids: returns (first_walker, n_local)."""
    base, rem = divmod(n_total, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


@dataclass(frozen=True)
class EstimatorLayout:
    n_params: int
    n_other: int

    @property
    def size(self):
        P = self.n_params
        return P * P + 3 * P + 2 + self.n_other + 3

    def slices(self):
        P, n = self.n_params, self.n_other
        o = 0
        out = {}
        for name, ln in (("S", P * P), ("OER", P), ("OEI", P), ("O", P), ("e_r", 1), ("e_i", 1), ("other", n),
                         ("n_acceptances", 1), ("n_trials", 1), ("n_samples", 1)):
            out[name] = slice(o, o + ln)
            o += ln
        return out

    def pack(self, S, OER, OEI, O, e_r, e_i, other, n_acc, n_trials, n_samples):
        buf = np.zeros(self.size)
        s = self.slices()
        buf[s["S"]] = np.asarray(S).ravel()
        buf[s["OER"]], buf[s["OEI"]], buf[s["O"]] = OER, OEI, O
        buf[s["e_r"]], buf[s["e_i"]] = e_r, e_i
        buf[s["other"]][:len(other)] = other
        buf[s["n_acceptances"]], buf[s["n_trials"]], buf[s["n_samples"]] = n_acc, n_trials, n_samples
        return buf

    def averages(self, buf):
        """Sums -> the averages the reference's root rank holds after ReduceToAverage."""
        s = self.slices()
        n = float(buf[s["n_samples"]][0])
        P = self.n_params
        return dict(localOperatorsMatrix=buf[s["S"]].reshape(P, P) / n, localOperatorlocalEnergyR=buf[s["OER"]] / n,
                    localOperatorlocalEnergyI=buf[s["OEI"]] / n, localOperators=buf[s["O"]] / n,
                    localEnergyR=float(buf[s["e_r"]][0]) / n, localEnergyI=float(buf[s["e_i"]][0]) / n,
                    otherExpectationValues=buf[s["other"]] / n, nAcceptances=int(round(buf[s["n_acceptances"]][0])),
                    nTrials=int(round(buf[s["n_trials"]][0])), nSamples=int(round(n)))


def allreduce_sum(buf, group=None):
    """Sum the packed buffer over the process group (gloo on CPU, nccl on GPU tensors)."""
    import torch
    import torch.distributed as dist

    t = torch.from_numpy(np.ascontiguousarray(buf))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.numpy()
