"""CPU check (no GPU) of the index arithmetic of the opt-in fused tile kernels (slate_b200/csrc/potrf_tile_fused.cu):
scratch/emulate_fused.py transcribes the loaders, the DMMA / FP32 product micro-kernels and the three block algorithms
thread by thread into numpy and compares them with dense references (ragged sizes included)."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fused_kernels_index_arithmetic_emulation():
    spec = importlib.util.spec_from_file_location("emulate_fused", os.path.join(ROOT, "scratch", "emulate_fused.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.main()


def test_panel_ll_protocol_emulation():
    """scratch/emulate_panel_ll.py: the tagged-slot exchange of getrf_base_ll_kernel (double-buffered by column parity, no
    barrier) run by G threads with shuffled interleavings: no deadlock, pivots / factors of plain partial pivoting."""
    spec = importlib.util.spec_from_file_location("emulate_panel_ll", os.path.join(ROOT, "scratch", "emulate_panel_ll.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.main()


def test_streaming_potrf_schedule_model():
    """scratch/emulate_stream_potrf.py: the stream / event schedule of potrf with streaming input (chunks, catch-up updates)
    executed in random event-respecting interleavings gives bitwise the right-looking factor and never touches a chunk
    before its arrival."""
    spec = importlib.util.spec_from_file_location("emulate_stream_potrf", os.path.join(ROOT, "scratch", "emulate_stream_potrf.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.main()


def test_fused_tile_flag_protocol_emulation():
    """scratch/emulate_tile_flags.py: the rowcnt / diagf release-acquire pipeline of potrf_tile_fused_kernel with one thread
    per CTA started in shuffled order: no deadlock, LAPACK's factor, first failing minor, FAILED propagation."""
    spec = importlib.util.spec_from_file_location("emulate_tile_flags", os.path.join(ROOT, "scratch", "emulate_tile_flags.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.main()
