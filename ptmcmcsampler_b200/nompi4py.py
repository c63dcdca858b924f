"""Single-process stand-in for ``mpi4py.MPI``: a world of one rank.

The engine keeps every temperature on the GPU, so there is no rank-to-rank traffic to carry; this
module only keeps code written against the reference working (``comm=nompi4py.COMM_WORLD`` as in ref
tests/test_simple.py:100-102, and the communicator surface listed in SURVEY.md section 2b).
"""


class MPIDummy(object):
    """Communicator of a world with a single member.  Point-to-point calls have no peer and do
    nothing; collectives return what rank 0 of a one-rank world would receive."""

    rank, size = 0, 1

    # ---- identity
    def Get_size(self):
        return self.size

    def Get_rank(self):
        return self.rank

    # ---- collectives over one rank
    def gather(self, sendobj, root=0, **kw):
        """Everything gathered from the world: the caller's own contribution."""
        return [sendobj]

    def scatter(self, sendobj, root=0, **kw):
        """The slice of ``sendobj`` addressed to rank 0 (``None`` scatters nothing)."""
        return sendobj[0] if sendobj is not None else None

    def bcast(self, obj, root=0, **kw):
        return obj

    def barrier(self):
        return None

    # ---- point to point: nobody to talk to
    def _no_peer(self, *args, **kw):
        return None

    send = recv = Iprobe = _no_peer


COMM_WORLD = MPIDummy()
