"""Shared helpers of the test-suite."""
import importlib
import os
import sys
import threading

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

    threads = [threading.Thread(target=body, args=(r,), daemon=True) for r in range(nranks)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout)
    alive = [t for t in threads if t.is_alive()]
    for e in errors:
        if e is not None:
            raise e
    assert not alive, "rank threads hung"
    group.destroy()
    return results
