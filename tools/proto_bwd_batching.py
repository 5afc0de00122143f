"""CPU model of render_bwd_kernel's staging control flow (csrc/render_bwd.cu, kHits path): the hit bytes of a tile's list are
scanned kBwdChunk positions at a time in DESCENDING list order, reachable positions (byte != 0) queue up in s_pos, and a batch
of at most kBwdBatch of them is gathered and replayed whenever kBwdBatch wait or the list is exhausted; a remainder shorter than
a batch moves to the front of the queue before the next scan.  `emulate` returns the batches in the order the kernel replays
them and the largest queue index it ever writes (capacity check: s_pos has kBwdChunk + kBwdBatch entries).
tests/test_host_logic.py checks it against the definition: the concatenated batches are the reachable positions, descending."""


def emulate(hit_bytes, n, chunk=256, batch=128, threads=64):
    per = chunk // threads
    chunks = (n + chunk - 1) // chunk
    queue = [None] * (chunk + batch)
    have = off = next_chunk = 0
    high_water = -1
    batches = []
    while True:
        if have < batch and next_chunk < chunks:
            if off > 0:                                   # remainder to the front
                moved = queue[off:off + have]
                queue[:have] = moved
                off = 0
            slot = have
            for tid in range(threads):                    # thread-major order == descending list position
                first_pos = n - 1 - (next_chunk * chunk + tid * per)
                for q in range(per):
                    pos = first_pos - q
                    if pos >= 0 and hit_bytes[pos] != 0:
                        queue[slot] = pos
                        high_water = max(high_water, slot)
                        slot += 1
            have = slot
            next_chunk += 1
            continue
        if have == 0:
            break
        total = min(batch, have)
        batches.append(queue[off:off + total])
        off += total
        have -= total
    return batches, high_water
