"""The engine's launch plan as host arithmetic (msm_plan_describe, include/msm_b200.h): window choice, sub-batches of
the pipelined scalar upload, task groups of many-task rows, slice length of the bucket kernel in whole waves.

No GPU is needed (and none is used): the plan is what make_plan (csrc/engine_impl.cuh) computes before any launch.
Without a device the bucket kernel's occupancy defaults to 4 blocks of 128 threads on each of 148 SMs.
"""
import pytest

WAVE = 148 * 4 * 128  # threads of one wave of k_accumulate with the CPU-side defaults


@pytest.fixture(scope="module")
def plan(engine):
    return engine.describe_plan


def test_headline_shape(plan):
    """BASELINE.json configs[2]: 2^24 BN254 points on the c = 22 window table."""
    p = plan(0, 1 << 24, table_window_bits=22)
    assert (p["window_bits"], p["num_windows"], p["buckets"]) == (22, 12, 1 << 21)
    assert p["digits_max"] == 12 << 24 and p["sort_mode"] == 2 and p["sub_batches"] == 1
    # the slice length is rounded UP to whole waves: 8 waves, never a block more (S = 332 made 4738 blocks = 8 waves + 2)
    assert p["wave_slices"] == WAVE and p["waves"] == 8  # BN254: 128 registers, 4 blocks per SM on the device too
    assert p["slice_len"] == 333
    assert p["slices"] <= 8 * WAVE and (p["slices"] + 127) // 128 == 4724


@pytest.mark.parametrize("curve,log_n,table_c", [(0, 20, 17), (0, 21, 20), (0, 23, 22), (1, 22, 20), (0, 22, 0), (1, 19, 0),
                                                   (0, 16, 0), (2, 20, 0), (3, 18, 0)])
def test_grid_is_whole_waves_rounded_down(plan, curve, log_n, table_c):
    p = plan(curve, 1 << log_n, table_window_bits=table_c)
    if p["waves"]:  # 0: the slice length was imposed (minimum of 8 digits, or an eighth of the average bucket)
        # 5 to 8 waves of 592 blocks' worth of threads, whatever the kernel's real blocks per SM are (on a device:
        # 3 for BLS12-381, 2 for the Fq2 fields, so more and smaller waves)
        assert 0.6 * 8 * WAVE <= p["waves"] * p["wave_slices"] <= 1.1 * 8 * WAVE
        assert p["slices"] <= p["waves"] * p["wave_slices"]
        # and not a whole wave short either
        assert p["slices"] > (p["waves"] - 1) * p["wave_slices"]
    assert 8 <= p["slice_len"] <= 1024
    assert p["slices"] == -(-p["digits_max"] // p["slice_len"])


def test_short_rows_take_fewer_longer_slices(plan):
    p = plan(0, 1 << 21, table_window_bits=20)  # 13 windows: 27.3 M digits
    assert p["num_windows"] == 13 and p["waves"] == 5 and p["slice_len"] == 72


@pytest.mark.parametrize("n_sub,growth", [(4, 2.0), (3, 3.0), (2, 2.0), (8, 1.0), (5, 1.5)])
def test_sub_batches_of_one_msm(plan, n_sub, growth):
    """Parts of one MSM: contiguous, covering the row, sizes growing by the factor asked for."""
    n = (1 << 24) + 12345
    p = plan(0, n, table_window_bits=22, sub_batches=n_sub, growth=growth)
    assert p["sub_batches"] == n_sub and p["by_task"] == 0
    first = p["sub_first"]
    assert first[0] == 0 and first[n_sub] == n and all(f == n for f in first[n_sub:])
    sizes = [first[k + 1] - first[k] for k in range(n_sub)]
    assert all(s > 0 for s in sizes)
    for a, b in zip(sizes, sizes[1:]):
        assert abs(b / a - growth) < 0.01
    # every sub-batch's digits fit the slices the plan reserves for the longest one
    assert p["slice_len"] * p["waves"] * p["wave_slices"] >= max(sizes) * p["num_windows"]


def test_first_upload_of_three_is_a_thirteenth(plan):
    p = plan(0, 13 << 20, table_window_bits=22, sub_batches=3, growth=3.0)
    assert p["sub_first"][:4] == [0, 1 << 20, 4 << 20, 13 << 20]


@pytest.mark.parametrize("num_chunks,n_sub", [(1024, 3), (1024, 4), (37, 3), (37, 8), (16, 8), (17, 4)])
def test_task_groups_are_whole_tasks(plan, num_chunks, n_sub):
    """A many-task row splits into groups of whole tasks (at least one each), sizes doubling; the dropped tail of the row
    (ag-build/cl/multiexp.cl:235) is in no group."""
    chunk_len = 4099
    n = num_chunks * chunk_len + 7
    p = plan(0, n, num_chunks=num_chunks, sub_batches=n_sub)
    assert p["sub_batches"] == n_sub and p["by_task"] == 1
    first = p["sub_first"]
    assert first[0] == 0 and first[n_sub] == num_chunks * chunk_len
    tasks = [(first[k + 1] - first[k]) for k in range(n_sub)]
    assert all(t > 0 and t % chunk_len == 0 for t in tasks)
    counts = [t // chunk_len for t in tasks]
    assert sum(counts) == num_chunks
    if num_chunks >= 64:
        for a, b in zip(counts, counts[1:]):
            assert 1.8 < b / a < 2.2


def test_groups_need_two_tasks_each_and_one_line(plan):
    assert plan(0, 1 << 20, num_chunks=7, sub_batches=4)["sub_batches"] == 1      # fewer than 2 tasks per group
    assert plan(0, 1 << 20, n_lines=10, num_chunks=2048, sub_batches=4)["sub_batches"] == 1  # several lines of bases
    assert plan(0, 1 << 20, n_lines=3, sub_batches=4)["sub_batches"] == 1


def test_window_choice_follows_the_cost_model(plan):
    """No table: the window balances 10 products per digit against the measured reduction cost per bucket."""
    assert plan(0, 1 << 24)["window_bits"] == 17 and plan(0, 1 << 24)["num_windows"] == 15
    # 1024 tasks of 2^12 points: per-window bucket sets of c = 8 .. 9; on a chunked table one set per task, c = 12
    p = plan(0, 1 << 22, num_chunks=1024)
    assert p["window_bits"] in (8, 9) and p["buckets"] == 1024 * p["num_windows"] << (p["window_bits"] - 1)
    t = plan(0, 1 << 22, num_chunks=1024, table_window_bits=12)
    assert t["num_windows"] == 22 and t["buckets"] == 1024 << 11
    # BLS12-381 scalars are 255 bits
    assert plan(1, 1 << 22, table_window_bits=20)["num_windows"] == 13


def test_invalid_shapes(engine):
    for args in [dict(n_scalars=0), dict(n_scalars=1 << 31), dict(n_scalars=100, num_chunks=101),
                 dict(n_scalars=1 << 20, n_lines=65536), dict(n_scalars=100, sub_batches=9)]:
        with pytest.raises(engine.CudaError):
            engine.describe_plan(0, **args)
    with pytest.raises(engine.CudaError):
        engine.describe_plan(7, 100)


def test_pipeline_shape_before_anything_is_measured(engine):
    """First call of a shape: short rows in one piece, one MSM in 2 parts from 2^20 scalars and 4 from 2^23 (sizes
    doubling), a many-task row in 3 task groups, several lines of bases in one piece."""
    ps = engine.pipeline_shape
    assert ps((1 << 20) - 1) == (1, 2.0)
    assert ps(1 << 20) == (2, 2.0) and ps((1 << 23) - 1) == (2, 2.0)
    assert ps(1 << 23) == (4, 2.0) and ps(1 << 24) == (4, 2.0)
    assert ps(1 << 22, num_chunks=1024) == (3, 2.0)
    assert ps(1 << 22, num_chunks=8) == (1, 2.0)                   # too few tasks for groups
    assert ps(1 << 19, num_chunks=128) == (1, 2.0)                 # an eighth of the batched row on one of 8 GPUs
    assert ps(1 << 21, n_lines=10, num_chunks=2048) == (1, 2.0)    # the AMT shape: 10 lines share the scalar row


def test_pipeline_shape_follows_the_measured_speeds(engine):
    """From the second call on the growth is 0.85 x device time / upload time, clamped to 1.5 .. 3; three parts when it
    reaches 2.5 (from 2^23 scalars), four otherwise."""
    ps = engine.pipeline_shape
    # one B200 with the PCIe link to itself: 55 GB/s, 2^24 points in 32.7 ms -> upload 9.76 ms, ratio 3.35
    n, g = ps(1 << 24, h2d_gbs=55.0, device_ms=32.7)
    assert n == 3 and abs(g - 0.85 * 32.7 / (536.870912 / 55.0)) < 1e-5 and 2.8 < g < 2.9
    # eight ranks on one host: 23 GB/s to a 2^21-point shard that takes 5.1 ms on the device -> 0.85 x 1.75 = 1.49,
    # held at the floor of 1.5; two parts
    assert ps(1 << 21, h2d_gbs=23.0, device_ms=5.1) == (2, 1.5)
    n, g = ps(1 << 21, h2d_gbs=36.0, device_ms=5.1)  # the four GPUs of that box with the faster links
    assert n == 2 and abs(g - 0.85 * 5.1 / (67.108864 / 36.0)) < 1e-5 and 2.3 < g < 2.4
    # a slow link under a large row: growth stays at the floor, four parts
    assert ps(1 << 24, h2d_gbs=12.0, device_ms=32.7) == (4, 1.5)
    # a very fast link: capped at 3
    assert ps(1 << 24, h2d_gbs=400.0, device_ms=32.7) == (3, 3.0)
    # only one of the two speeds known: as if nothing had been measured
    assert ps(1 << 24, h2d_gbs=55.0) == (4, 2.0) and ps(1 << 24, device_ms=32.7) == (4, 2.0)
    # task groups do not adapt
    assert ps(1 << 22, num_chunks=1024, h2d_gbs=55.0, device_ms=16.3) == (3, 2.0)
    with pytest.raises(engine.CudaError):
        ps(0)


def test_plan_invariants_over_random_shapes(engine):
    """Whatever the shape: sub-batches are contiguous and cover exactly the scalars used, task groups are whole tasks,
    the slices cover every digit, and a grid sized in waves never spills into one more."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    @settings(max_examples=300, deadline=None)
    @given(curve=st.integers(0, 3), log_l=st.integers(4, 25), jitter=st.integers(0, 1000), chunks_log=st.integers(0, 12),
           chunk_jitter=st.integers(0, 3), n_lines=st.sampled_from([1, 1, 1, 2, 10]), n_sub=st.integers(1, 8),
           growth=st.sampled_from([1.0, 1.5, 2.0, 2.9, 3.0]), table=st.booleans())
    def check(curve, log_l, jitter, chunks_log, chunk_jitter, n_lines, n_sub, growth, table):
        L = (1 << log_l) + jitter
        num_chunks = min((1 << chunks_log) + chunk_jitter, L)
        table_c = 0
        if table:
            table_c = min(22, max(8, log_l - chunks_log))
        try:
            p = engine.describe_plan(curve, L, n_lines=n_lines, num_chunks=num_chunks, table_window_bits=table_c,
                                     sub_batches=n_sub, growth=growth)
        except engine.CudaError:
            return  # too many buckets / digits for 32-bit indices: rejected, not planned
        chunk_len = L // num_chunks
        used = chunk_len * num_chunks
        k = p["sub_batches"]
        first = p["sub_first"]
        assert 1 <= k <= n_sub and first[0] == 0 and first[k] == used
        assert all(first[i] <= first[i + 1] for i in range(k))
        if p["by_task"]:
            assert n_lines == 1 and num_chunks >= 2 * k
            assert all(first[i] < first[i + 1] and first[i] % chunk_len == 0 for i in range(k))
        if n_lines != 1:
            assert k == 1
        assert p["digits_max"] == used * p["num_windows"]
        assert p["slices"] * p["slice_len"] >= p["digits_max"] and 8 <= p["slice_len"] <= 1024
        sets = 1 if table_c else p["num_windows"]
        assert p["buckets"] == num_chunks * sets << (p["window_bits"] - 1)
        if p["waves"]:
            longest = max(first[i + 1] - first[i] for i in range(k))
            assert -(-longest * p["num_windows"] // p["slice_len"]) <= p["waves"] * p["wave_slices"]

    check()
