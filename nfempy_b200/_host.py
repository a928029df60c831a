"""Host-side thread pool for the memory-bound numpy passes of the drop-in call (state staging, ``residual += rhs``, the
coefficient callables over 10^7 quadrature points): numpy releases the GIL inside its loops, so a few threads over chunks
cost a fraction of one single-threaded sweep.  ``NFA_HOST_THREADS=1`` switches it off."""
import os

_POOL = None


def n_threads():
    env = os.environ.get("NFA_HOST_THREADS")
    if env:
        return max(1, int(env))
    ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))  # torchrun: one process per GPU shares the host cores
    return max(1, min(8, (os.cpu_count() or 4) // (2 * ranks)))


def pool():
    """The shared executor, or ``None`` when threading is switched off."""
    global _POOL
    if n_threads() <= 1:
        return None
    if _POOL is None:
        from concurrent.futures import ThreadPoolExecutor

        _POOL = ThreadPoolExecutor(max_workers=n_threads())
    return _POOL


def parts(n, chunk):
    return [(a, min(n, a + chunk)) for a in range(0, n, chunk)]
