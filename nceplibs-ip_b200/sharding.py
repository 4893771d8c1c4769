"""Field-batch sharding across GPUs.

The km fields of one ipolates / ipolatev call are independent given the plan (ibo(k) and the pole fix are per
field), so rank r of G interpolates a contiguous slice of fields and no data-path collective is needed.  Vector
pairs (u_k, v_k) share an index and therefore stay together.
"""


def field_slice(km, world, rank):
    """Contiguous, balanced partition of range(km): the first km % world ranks get one extra field."""
    if not (0 <= rank < world) or km < 0:
        raise ValueError((km, world, rank))
    base, extra = divmod(km, world)
    k0 = rank * base + min(rank, extra)
    return k0, k0 + base + (1 if rank < extra else 0)


def weak_slice(km_per_gpu, world, rank):
    """bench.py's weak-scaling layout: every rank owns km_per_gpu fields of a (world*km_per_gpu)-field batch."""
    return rank * km_per_gpu, (rank + 1) * km_per_gpu
