"""Shared-memory bank-conflict model of the FFT kernel's exchange / staging accesses.
Counts wavefronts per warp instruction (ideal = bytes/128) for every access pattern of a config."""
import itertools, sys

PLANS = {32: (8, [8, 4]), 64: (8, [8, 8]), 128: (16, [16, 8]), 256: (16, [16, 16]), 512: (8, [8, 8, 8]), 1024: (16, [16, 16, 4]),
         2048: (16, [16, 16, 8]), 4096: (16, [16, 16, 16])}


def wavefronts(addrs, eb):
    """addrs: byte addresses of the 32 lanes (None = inactive); eb = bytes per lane access"""
    group = {16: 8, 8: 16, 4: 32}[eb]
    total = 0
    for g0 in range(0, 32, group):
        lanes = [a for a in addrs[g0:g0 + group] if a is not None]
        banks = {}
        for a in lanes:
            for w in range(eb // 4):
                word = a // 4 + w
                banks.setdefault(word % 32, set()).add(word)
        total += max((len(v) for v in banks.values()), default=0)
    return total


def run(N, eb, TX, PADK=None):
    E, radices = PLANS[N]
    PADK = PADK or radices[0]
    T = N // E
    q = 128 // eb
    raw = (N - 1) + (N - 1) // PADK + 1
    step = max(1, q // TX)
    LS = ((raw + q - 1 - step) // q * q + step) if TX > 1 else raw
    pad = lambda p: p + p // PADK
    threads = TX * T * max(1, 256 // (TX * T))
    out = {}

    def measure(name, addr_fn, ident):
        worst, tot, n = 0, 0, 0
        for w0 in range(0, min(threads, 256), 32):
            for s in range(E):
                addrs = []
                for t in range(w0, w0 + 32):
                    if ident == "tile":
                        tx, j, ly = t % TX, (t // TX) % T, t // (TX * T)
                    else:
                        j, tx, ly = t % T, (t // T) % TX, t // (TX * T)
                    p = addr_fn(j, s)
                    addrs.append(None if p is None else ((ly * TX + tx) * LS + pad(p)) * eb)
                wf = wavefronts(addrs, eb)
                ideal = max(1, sum(a is not None for a in addrs) * eb // 128)
                worst = max(worst, wf / ideal)
                tot += wf
                n += ideal
        out[name] = (tot / n, worst)

    measure("read j+T*s (tile lanes)", lambda j, s: j + T * s, "tile")
    measure("stage j+T*s (line lanes)", lambda j, s: j + T * s, "line")
    ns = 1
    for pi, R in enumerate(radices[:-1]):
        NB = E // R

        def scat(j, s, R=R, NB=NB, ns=ns):
            u, r = s % NB, s // NB
            jj = j + T * u
            return (jj // ns) * (ns * R) + jj % ns + r * ns
        measure(f"scatter pass {pi} (R={R}, Ns={ns})", scat, "tile")
        ns *= R
    measure("r2c partner (N-k)%N", lambda j, s: (N - (j + T * s)) % N if s <= E // 2 and j + T * s <= N // 2 else None, "tile")
    return LS, out


if __name__ == "__main__":
    for N, eb, TX in ((1024, 16, 4), (1024, 16, 1), (512, 16, 4), (2048, 16, 4), (256, 16, 4), (1024, 8, 8), (2048, 8, 8), (512, 8, 8), (1024, 16, 8)):
        LS, out = run(N, eb, TX)
        print(f"N={N} elem={eb}B TX={TX} line_sm={LS}")
        for k, (avg, worst) in out.items():
            print(f"   {k:34s} avg x{avg:4.2f} worst x{worst:4.2f}")
