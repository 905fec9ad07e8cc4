"""In-process message fabric for the protocol tests: n mailboxes, optional
seeded random delivery delay (reordering), like the reference's TestRouter
(tests/fixtures.py:116-141) -- written for these tests, no reference code."""

import asyncio
import random


class SimNet:
    def __init__(self, n, max_delay=0.0, seed=0):
        self.n = n
        self.boxes = [asyncio.Queue() for _ in range(n)]
        self.rng = random.Random(seed)
        self.max_delay = max_delay
        self.sends = [self._make_send(i) for i in range(n)]
        self.recvs = [self._make_recv(i) for i in range(n)]

    def _make_send(self, me):
        def send(dest, message):
            if self.max_delay > 0:
                delay = self.rng.random() * self.max_delay
                asyncio.get_event_loop().call_later(delay, self.boxes[dest].put_nowait, (me, message))
            else:
                self.boxes[dest].put_nowait((me, message))

        return send

    def _make_recv(self, me):
        async def recv():
            return await self.boxes[me].get()

        return recv


def run(coro):
    loop = asyncio.new_event_loop()
    try:
        return loop.run_until_complete(coro)
    finally:
        loop.close()
