"""Single-process stand-in for ``mpi4py.MPI`` with the surface of the reference's
``PTMCMCSampler/nompi4py.py``: rank 0 of a world of size 1.

The engine keeps every temperature on the GPU, so no rank-to-rank traffic exists; this module is
here so that code written against the reference (``comm=nompi4py.COMM_WORLD``,
ref tests/test_simple.py:100-102) keeps working.
"""


class MPIDummy(object):
    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def barrier(self):
        return None

    def send(self, obj, dest=1, tag=55):
        return None

    def recv(self, source=1, tag=55):
        return None

    def Iprobe(self, source=1, tag=55):
        return None

    def scatter(self, sendobj, **kwargs):
        return None if sendobj is None else sendobj[0]

    def bcast(self, obj, **kwargs):
        return obj

    def gather(self, sendobj, **kwargs):
        return [sendobj]


COMM_WORLD = MPIDummy()
