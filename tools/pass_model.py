"""development aid: list-scheduling model of the action-reaction launch (engine.cu: launch_pair_sym / sym_plan).  Each pass launches
(rows of the j-side buffer) x (j-chunks) CTAs in chunk-major order onto `slots` resident-CTA slots; a CTA costs the number of active tiles of
its (i-block, chunk).  Prints (i-blocks, chunks, busy fraction = sum of CTA costs / (makespan x slots), summed over the passes).
Used to choose the chunk-count rule of steps_b200_sym_chunk_target."""
import heapq, sys
def model(n, ib, tj, n_chunks_target, rows_per_pass, slots, min_tpc=1):
    n_tiles = (n + tj - 1)//tj
    tpb = ib//tj
    nb = (n + ib - 1)//ib
    tpc = max(min_tpc, (n_tiles + n_chunks_target - 1)//n_chunks_target)
    n_chunks = (n_tiles + tpc - 1)//tpc
    total_busy = 0.0; total_span = 0.0
    for b0 in range(0, nb, rows_per_pass):
        blocks = range(b0, min(nb, b0 + rows_per_pass))
        # chunk-major order
        works = []
        for jc in range(n_chunks):
            c0, c1 = jc*tpc, min((jc+1)*tpc, n_tiles)
            for b in blocks:
                dlo, dhi = b*tpb, min((b+1)*tpb, n_tiles)
                # diag tiles cost 1 unit/tile (one-sided: R*128 pairs each), sym tiles cost 1 unit/tile too (same pair count but both sides)
                lo = max(dlo, c0); hi = c1
                w = max(0, hi - lo)
                works.append(w + 0.02 if w > 0 else 0.002)
        h = [0.0]*slots
        heapq.heapify(h)
        for w in works:
            t = heapq.heappop(h); heapq.heappush(h, t + w)
        span = max(h); busy = sum(works)
        total_busy += busy; total_span += span*slots
    return nb, n_chunks, total_busy/total_span
for name, args in [("C2 f64", (2_000_000, 768, 128, 56, 341, 296)),
                   ("C5 f32 16GB", (16_777_216, 1024, 128, 56, 79, 592)),
                   ("C5 f32 48GB", (16_777_216, 1024, 128, 56, 238, 592)),
                   ("C5 f32 16GB 224 chunks", (16_777_216, 1024, 128, 224, 79, 592)),
                   ("C5 f32 48GB 112 chunks", (16_777_216, 1024, 128, 112, 238, 592)),
                   ("C4 s1r2 16GB", (4_194_304, 384, 128, 56, 160, 592)),
                   ("C3 t3 gen 16GB", (2_097_152, 256, 128, 56, 320, 444)),
                   ("2M f32", (2_000_000, 1024, 128, 56, 1954, 592))]:
    print(name, model(*args))
