"""An in-process stand-in for ``torch.distributed`` (CPU tensors, one THREAD per rank): exactly the
calls photon_weave_b200/sharded.py makes -- all_reduce (sum), all_to_all_single (even or split),
all_gather, batch_isend_irecv / P2POp / isend / irecv -- so that many randomised multi-rank cases
run in milliseconds.  The gloo tests (tests/test_sharded_gloo.py) cover the real backend."""
import queue
import threading


class _Req:
    def wait(self):
        return None


class _Shared:
    def __init__(self, world):
        self.world = world
        self.barrier = threading.Barrier(world)
        self.slots = [None] * world
        self.mail = {(s, d): queue.Queue() for s in range(world) for d in range(world)}


class P2POp:
    def __init__(self, op, tensor, peer):
        self.op, self.tensor, self.peer = op, tensor, peer


class RankDist:
    """The ``dist`` object of one rank."""

    P2POp = P2POp

    def __init__(self, shared, rank):
        self.s, self.rank = shared, rank

    # markers compared by identity in batch_isend_irecv
    @staticmethod
    def isend(*a, **k):
        raise NotImplementedError

    @staticmethod
    def irecv(*a, **k):
        raise NotImplementedError

    def _exchange(self, item):
        self.s.slots[self.rank] = item
        self.s.barrier.wait()
        got = list(self.s.slots)
        self.s.barrier.wait()
        return got

    def all_reduce(self, t, op=None):
        parts = self._exchange(t.clone())
        acc = parts[0].clone()
        for p in parts[1:]:
            acc += p
        t.copy_(acc)

    def all_gather(self, outs, t):
        for o, p in zip(outs, self._exchange(t.clone())):
            o.copy_(p)

    def all_to_all_single(self, out, inp, output_split_sizes=None, input_split_sizes=None):
        W = self.s.world
        n = inp.shape[0]
        ins = input_split_sizes or [n // W] * W
        outs = output_split_sizes or [out.shape[0] // W] * W
        parts = self._exchange((inp.clone(), ins))
        pos = 0
        for src in range(W):
            data, splits = parts[src]
            start = sum(splits[:self.rank])
            piece = data[start:start + splits[self.rank]]
            assert piece.shape[0] == outs[src]
            out[pos:pos + outs[src]].copy_(piece)
            pos += outs[src]

    def batch_isend_irecv(self, ops):
        for o in ops:
            if o.op is RankDist.isend or getattr(o.op, "__name__", "") == "isend":
                self.s.mail[(self.rank, o.peer)].put(o.tensor.clone())
        for o in ops:
            if o.op is RankDist.irecv or getattr(o.op, "__name__", "") == "irecv":
                o.tensor.copy_(self.s.mail[(o.peer, self.rank)].get(timeout=60))
        return [_Req() for _ in ops]

    def barrier(self):
        self.s.barrier.wait()


def run_ranks(world, fn, *args):
    """Run ``fn(rank, world, dist, *args)`` on ``world`` threads; returns the per-rank results and
    re-raises the first failure."""
    shared = _Shared(world)
    results, errors = [None] * world, []

    def body(r):
        try:
            results[r] = fn(r, world, RankDist(shared, r), *args)
        except BaseException as exc:  # noqa: BLE001 -- reported to the caller
            errors.append((r, exc))
            shared.barrier.abort()

    threads = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(120)
    real = [e for e in errors if not isinstance(e[1], threading.BrokenBarrierError)]
    if real or errors:
        raise (real or errors)[0][1]
    return results
