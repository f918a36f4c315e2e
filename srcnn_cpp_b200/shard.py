"""Host-side sharding plans for the multi-GPU path: the hot path has no exchange step, so
ranks take disjoint frames or disjoint output-row bands and never talk on the data path.  Pure
arithmetic -- no torch, no CUDA -- so the plans are unit-tested on CPU (tests/test_sharding_gloo.py)."""


def frames_for_rank(n_frames, rank, world):
    """Frame f goes to GPU f mod world (BASELINE configs[2]: 'frame-sharded across 1/2/4/8 B200')."""
    return list(range(rank, n_frames, world))


def band_edges(out_rows, n_bands):
    """n_bands contiguous output-row bands covering [0, out_rows): band i = [edges[i], edges[i+1])."""
    n_bands = max(1, min(n_bands, out_rows))
    return [out_rows * i // n_bands for i in range(n_bands + 1)]


def bands_for_rank(out_rows, rank, world, bands_per_rank=1):
    """Contiguous bands of one rank when an image is split into world*bands_per_rank row bands
    (BASELINE configs[3]: 'split into row bands with 6-px halo across 8 B200')."""
    edges = band_edges(out_rows, world * bands_per_rank)
    nb = len(edges) - 1
    mine = [i for i in range(nb) if i * world // nb == rank] if nb >= world else ([rank] if rank < nb else [])
    return [(edges[i], edges[i + 1]) for i in mine]


def band_src_rows_py(h, scale, r0, r1, taps_ofs):
    """Source rows [s0, s1) that output rows [r0, r1) depend on, given the vertical tap offsets
    (floor of the source coordinate per output row).  Mirrors srcnn_band_src_rows in the C ABI:
    6-px halo in the upscaled-Y domain (4 conv1 + 2 conv3), then the 4-tap cubic footprint."""
    oh = len(taps_ofs)
    p0, p1 = max(r0 - 6, 0), min(r1 + 6, oh)
    s0 = min(max(int(taps_ofs[p0]) - 1, 0), h - 1)
    s1 = min(max(int(taps_ofs[p1 - 1]) + 2, 0), h - 1) + 1
    return s0, s1
