"""Single-process stand-in for the un-vendored ``dlp_mpi`` package so that the
unmodified reference (``/root/reference/pb_chime5``) imports in this container.
Only used by ``oracle/make_golden.py`` -- never by the product."""
IS_MASTER = True
MASTER = 0
RANK = 0
SIZE = 1


def barrier():
    pass


def bcast(obj, root=0):
    return obj


def split_managed(sequence, allow_single_worker=False, **kwargs):
    yield from sequence
