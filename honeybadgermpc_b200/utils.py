"""List reshapes and message plumbing used by ``batch_reconstruct``
(reference: honeybadgermpc/utils/misc.py:21-106).  Behaviour is the same; the
``@TypeCheck`` decorators of the reference are not reproduced."""

import asyncio
from collections import defaultdict


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
