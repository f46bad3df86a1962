"""dev: probe statistics of the per-warp visited table (csrc/search.cuh) for candidate table sizes: mean lockstep steps per
batch of 32 ids (the loop runs until the slowest lane is done) and the share of queries that exhaust a 15-entry probe window.
Model of the C2 workload at ef = 57: ~72 batches per query, ~23 valid ids per batch, about half of them seen before;
distinct ids per query (= evaluations) log-normal with mean 800 and p99 = 1.7 x the mean (what the counters of the C2 bench show)."""
import numpy as np


def run(T, nq=2000, seed=0):
    rng = np.random.default_rng(seed)
    steps, over = [], 0
    for _ in range(nq):
        total = int(800 * np.exp(rng.normal(-0.026, 0.228)))
        universe = rng.choice(1 << 21, size=total, replace=False)
        tab = np.full(T, -1, np.int64)
        seen = 0
        ovf = False
        while seen < total:
            new = universe[seen:seen + 12]
            seen += len(new)
            old = universe[rng.integers(0, max(seen - len(new), 1), 11)] if seen > 12 else universe[:0]
            batch = np.unique(np.concatenate([new, old]))
            home = (batch * 0x9E3779B1 % (1 << 21)) * T >> 21
            worst = 0
            for b, h in zip(batch, home):
                for d in range(16):
                    s = (h + d) % T
                    if tab[s] == b or tab[s] == -1:
                        if d < 15:
                            tab[s] = b
                        break
                else:
                    d = 15
                if d >= 15:
                    ovf = True
                worst = max(worst, min(d, 14) + 1)
            steps.append(worst)
        over += ovf
    return np.mean(steps), over / nq


if __name__ == "__main__":
    for T in (4096, 3840, 3584, 3072, 2560, 2048):
        s, o = run(T)
        print(f"T={T}: {s:.2f} lockstep steps per batch, {100 * o:.1f}% of the queries exhaust a probe window")
