"""Progress and cancel on the STREAMING path (DelayCallback.delay / DelayAbortedException, Common/DelayCallback.java:34-56;
ProgressDialog.java:205-221): a long pass goes out in segments; between segments the library calls the progress callback on the
calling thread and honours tdna_cancel.  TDNA_OPT_SEGMENT_ITEMS makes the segments short enough for a test-sized alignment."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    import taxondna_b200 as t
    from taxondna_b200 import synth
    data = synth.generate("C3", seed=11, genera=20, species=10, reps=10)  # 2,000 x 658
    ctx = t.Context(0)
    aln = t.Alignment(ctx, data["symbols"])
    aln.set_labels(data["species_id"], data["genus_id"], data["epithet_rank"], data["fullname_rank"])
    yield t, ctx, aln
    ctx.set_progress(None)
    ctx.set_option(t.OPT_SEGMENT_ITEMS, 0)
    ctx.set_option(t.OPT_COUNT_KERNEL, t.KERNEL_AUTO)
    ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_AUTO)
    aln.close()
    ctx.close()


@pytest.mark.parametrize("kernel,items", [("tensor", 8), ("popc", 64)])
@pytest.mark.parametrize("mode", [0, 1])
def test_segmented_pass_reports_progress_and_matches(setup, kernel, items, mode):
    t, ctx, aln = setup
    aln.set_config(mode, 300, True)
    ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_STREAM)
    ctx.set_option(t.OPT_COUNT_KERNEL, t.KERNEL_TENSOR if kernel == "tensor" else t.KERNEL_POPC)
    ctx.set_option(t.OPT_SEGMENT_ITEMS, 0)
    ctx.set_progress(None)
    whole = aln.best_match()
    calls = []
    ctx.set_progress(lambda done, total: calls.append((done, total)))
    ctx.set_option(t.OPT_SEGMENT_ITEMS, items)
    cut = aln.best_match()
    assert cut.tobytes() == whole.tobytes()
    assert len(calls) >= 2, calls
    assert all(0 < d <= tot for d, tot in calls)
    assert all(b[0] > a[0] for a, b in zip(calls, calls[1:]) if a[1] == b[1])


@pytest.mark.parametrize("kernel,items", [("tensor", 8), ("popc", 64)])
def test_cancel_mid_pass(setup, kernel, items):
    t, ctx, aln = setup
    aln.set_config(0, 300, True)
    ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_STREAM)
    ctx.set_option(t.OPT_COUNT_KERNEL, t.KERNEL_TENSOR if kernel == "tensor" else t.KERNEL_POPC)
    ctx.set_option(t.OPT_SEGMENT_ITEMS, 0)
    ctx.set_progress(None)
    whole = aln.best_match()
    seen = []

    def cb(done, total):  # the user presses Cancel while the second segment runs
        seen.append(done)
        if len(seen) == 2:
            ctx.cancel()

    ctx.set_progress(cb)
    ctx.set_option(t.OPT_SEGMENT_ITEMS, items)
    with pytest.raises(t.TdnaError) as ei:
        aln.best_match()
    assert ei.value.code == t.E_CANCELLED
    assert len(seen) == 2  # nothing ran past the boundary that saw the flag
    # the alignment stays usable, the flag is consumed
    ctx.set_progress(None)
    again = aln.best_match()
    assert again.tobytes() == whole.tobytes()
