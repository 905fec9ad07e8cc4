"""Message plumbing used by ``batch_reconstruct`` (reference:
honeybadgermpc/utils/misc.py:76-106).  The reference's list reshapes
(``chunk_data / transpose_lists / flatten_lists``, utils/misc.py:33-73) have no
counterpart here: ``batch_reconstruct`` does them as numpy reshapes of limb arrays."""

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
