"""Reference-equivalent GPU baseline: OUR transcription of the structure of the upstream rasterizer
(per-view launch sequence, num_rendered read-back, global radix sort, block-wide batches, per-pixel atomics), used by
bench.py and tests only.  Never imported by the product package."""
