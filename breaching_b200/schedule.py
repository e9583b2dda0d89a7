"""Step-size tables for the device-side optimiser.

The engine indexes a device-resident table ``lr[it]`` instead of stepping scheduler objects on the host every
iteration.  The closed forms reproduce what the reference obtains from
``attacks/auxiliaries/common.py:19-38`` (MultiStepLR with float milestones, CosineAnnealingLR, LambdaLR) wrapped in
``GradualWarmupScheduler`` (``common.py:74-162``); ``tests/test_schedule.py`` checks them entry by entry against
scheduler objects stepped by the reference's own ``optimizer_lookup``.
"""
import math


def lr_table(step_size, scheduler=None, warmup=0, max_iterations=10_000, n=None):
    """``table[it]`` = step size used by optimiser step ``it`` (0-based)."""
    T = int(max_iterations)
    n = T if n is None else int(n)
    base = float(step_size)
    warmup = int(warmup or 0)
    if scheduler == "step-lr":
        # milestones stay floats (24000 // 1.6 == 14999.0); an integer epoch e has passed milestone m iff e >= m
        milestones = [T // 2.667, T // 1.6, T // 1.142]

        def after(e):
            return base * (0.1 ** sum(1 for m in milestones if e >= m))
    elif scheduler == "cosine-decay":
        def after(e):
            return base * (1 + math.cos(math.pi * e / T)) / 2 if T > 0 else base
    elif scheduler == "linear":
        def after(e):
            return base * max(0.0, float(T - e) / float(max(1, T)))
    else:
        def after(e):
            return base

    table = []
    for it in range(n):
        if warmup > 0:
            # warm-up: lr = base * epoch / warmup, i.e. the very first step runs at lr = 0; the wrapped scheduler's
            # clock starts only once the warm-up is over (common.py:100-117,131-146)
            table.append(base * (float(it) / warmup) if it <= warmup else after(it - warmup - 1))
        else:
            table.append(after(it))
    return table
