"""Shared helpers of the test-suite."""
import importlib
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pkg():
    return importlib.import_module("2decomp-fft_b200")


def run_ranks(nranks, fn, timeout=300):
    """Run fn(rank, group) on `nranks` threads sharing one d2d group (thread-per-rank transport);
    returns the list of results, re-raises the first exception."""
    p = pkg()
    group = p.Group(nranks)
    results, errors = [None] * nranks, [None] * nranks

    def body(r):
        try:
            results[r] = fn(r, group)
        except BaseException as e:  # noqa
            errors[r] = e
            group.abort()  # the ranks waiting for this one in an exchange must fail too, not hang

    threads = [threading.Thread(target=body, args=(r,), daemon=True) for r in range(nranks)]
    for t in threads:
        t.start()
    deadline = time.monotonic() + timeout  # one deadline for the whole group, not one per thread
    for t in threads:
        t.join(max(0.0, deadline - time.monotonic()))
    alive = [t for t in threads if t.is_alive()]
    if alive:
        group.abort()
    first = [e for e in errors if e is not None and "rank group aborted" not in str(e)] or [e for e in errors if e is not None]
    if first:
        raise first[0]
    assert not alive, "rank threads hung"
    group.destroy()
    return results
