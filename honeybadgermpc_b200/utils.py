"""List reshapes and message plumbing used by ``batch_reconstruct``
(reference: honeybadgermpc/utils/misc.py:21-106).  Behaviour is the same; the
``@TypeCheck`` decorators of the reference are not reproduced."""

import asyncio
import contextlib
import gc
from collections import defaultdict


@contextlib.contextmanager
def gc_paused():
    """Keep the cyclic garbage collector from running inside a bulk conversion.

    Turning B limbs into B Python objects (or chunking B ints into lists) allocates tens of
    thousands of container objects in one go; every 700 of them the collector starts a pass,
    and its older generations walk every tracked object of the process.  None of the objects
    made here can be part of a cycle.  Measured on a 49 152-share open at n = 16: collections
    triggered by these conversions cost more than the conversions themselves.  Only used
    around synchronous sections (no ``await`` inside)."""
    was_enabled = gc.isenabled()
    gc.disable()
    try:
        yield
    finally:
        if was_enabled:
            gc.enable()


def wrap_send(tag, send):
    """send(dest, msg) -> send(dest, (tag, msg))   (utils/misc.py:21-30)"""

    def tagged(dest, message):
        send(dest, (tag, message))

    return tagged


def chunk_data(data, chunk_size, default=0):
    """[1,2,3,4,5], 2 -> [[1,2],[3,4],[5,0]]; the empty list gives one flat chunk
    of defaults, exactly like the reference (utils/misc.py:33-52)."""
    if not data:
        return [default] * chunk_size
    chunks = [list(data[i: i + chunk_size]) for i in range(0, len(data), chunk_size)]
    chunks[-1].extend([default] * (chunk_size - len(chunks[-1])))
    return chunks


def flatten_lists(lists):
    return [v for inner in lists for v in inner]


def transpose_lists(lists):
    """[[1,2,3],[4,5,6]] -> [[1,4],[2,5],[3,6]]   (utils/misc.py:67-73)"""
    width = len(lists[0])
    return [[row[i] for row in lists] for i in range(width)]


def subscribe_recv(recv):
    """Demultiplex ``recv() -> (sender, (tag, payload))`` into one queue per tag
    (utils/misc.py:76-106).  Returns (background task, subscribe(tag) -> getter)."""
    queues = defaultdict(asyncio.Queue)
    taken = set()

    async def pump():
        while True:
            sender, (tag, payload) = await recv()
            queues[tag].put_nowait((sender, payload))

    def subscribe(tag):
        assert tag not in taken
        taken.add(tag)
        return queues[tag].get

    return asyncio.create_task(pump()), subscribe
