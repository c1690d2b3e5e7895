"""Host-side logic of the streaming pipeline (no GPU): the chunk schedule covers the batch exactly."""
import pytest

from algebraist_b200.host_pipeline import chunk_schedule


@pytest.mark.parametrize("batch,chunk", [(128, 8), (128, 16), (128, 32), (5, 8), (16, 8), (33, 8), (1, 4), (7, 1)])
@pytest.mark.parametrize("ramp", [True, False])
def test_chunk_schedule_covers_batch(batch, chunk, ramp):
    sizes = chunk_schedule(batch, chunk, ramp)
    assert sum(sizes) == batch and all(1 <= m <= chunk for m in sizes)
    if not ramp:
        assert sizes[:-1] == [chunk] * (len(sizes) - 1)
    elif len(sizes) > 4 and chunk >= 4:
        assert sizes[0] == chunk // 4 and sizes[-1] == chunk // 4      # short first and last chunk
