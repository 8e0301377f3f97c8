"""Host-side logic that needs no GPU: parameter table, flow expansion, .flo reader, PSNR."""
import numpy as np
import pytest
import torch


def test_params_match_reference_table():
    import vnlb_b200
    p = vnlb_b200.default_params(20.)
    assert p["nSimilarPatches"] == [100, 60] and p["sizePatch"] == [7, 7] and p["sizePatchTime"] == [2, 2]
    assert p["sizeSearchWindow"] == [27, 27] and p["rank"] == [39, 39] and p["variThres"] == [2.7, 0.7]
    assert p["nstreams"] == [8, 18] and p["bsize"] == [128, 128] and p["gamma"] == [0.95, 0.2]
    a0 = vnlb_b200.get_args(p, 3, 0, "cuda:0")
    a1 = vnlb_b200.get_args(p, 3, 1, "cuda:0")
    assert a0.patch_shape == (1024, 100, 2, 3, 7, 7) and a1.bufs_shape == (2304, 60)
    assert a0.sigma2 == 400. and a0.sigmab2 == 400. and a1.sigmab2 == 0. and a1.thresh == 0.7
    assert a0.dist_chnls == 1 and a1.dist_chnls == 3 and a0.w_s == 27 and a0.nWt_f == 6
    ip = vnlb_b200.get_params(20., version="iphone")
    assert ip["sizeSearchWindow"] == [15, 15] and ip["sizePatchTime"] == [1, 2]
    with pytest.raises(ValueError):
        vnlb_b200.get_params(20., version="nope")


def test_expand_flows_and_errors():
    from vnlb_b200.utils import expand_flows
    ff = torch.arange(3 * 2 * 2 * 2, dtype=torch.float32).reshape(3, 2, 2, 2)
    bf = -ff
    f2, b2 = expand_flows(dict(fflow=ff, bflow=bf), 4)       # lib/vnlb/utils/utils.py:143-158
    assert f2.shape[0] == 4 and torch.equal(f2[3], ff[2]) and torch.equal(b2[0], bf[0]) and torch.equal(b2[1:], bf)
    with pytest.raises(ValueError):
        expand_flows(dict(fflow=ff, bflow=bf), 6)
    with pytest.raises(ValueError):
        expand_flows(dict(fflow=ff, bflow=bf[:2]), 4)


def test_read_flo(tmp_path):
    from vnlb_b200.utils import read_flo
    h, w = 3, 5
    data = np.random.RandomState(0).randn(h, w, 2).astype(np.float32)
    path = tmp_path / "a.flo"
    with open(path, "wb") as f:
        np.array([202021.25], np.float32).tofile(f)
        np.array([w, h], np.int32).tofile(f)
        data.tofile(f)
    flo = read_flo(str(path))
    assert flo.shape == (2, h, w) and np.array_equal(flo[0], data[..., 0]) and np.array_equal(flo[1], data[..., 1])


def test_psnr_formula():
    from vnlb_b200 import compute_psnrs
    a = np.zeros((2, 3, 4, 4), np.float32)
    b = np.full((2, 3, 4, 4), 25.5, np.float32)
    np.testing.assert_allclose(compute_psnrs(a, b), [20., 20.], atol=1e-5)


def test_denoise_refuses_to_run_without_cuda():
    import vnlb_b200
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError):
        vnlb_b200.denoise(np.zeros((2, 3, 16, 16), np.float32), 20.)
    from vnlb_b200 import color
    with pytest.raises(ValueError):
        color.rgb2yuv(torch.zeros(1, 3, 4, 4))


def test_synth_matches_oracle_generator():
    from vnlb_b200 import synth
    from oracle import vnlb_oracle as orc
    assert np.array_equal(synth.synth_video(3, 24, 32), orc.synth_video(3, 24, 32))
    c = synth.synth_video(2, 16, 16)
    assert np.array_equal(synth.add_noise(c, 20.), orc.add_noise(c, 20.))


def test_video_io_roundtrip_and_report(tmp_path):
    from vnlb_b200 import video_io
    rs = np.random.RandomState(0)
    vid = rs.randint(0, 256, (3, 3, 10, 12)).astype(np.float32)
    video_io.save_video_sequence(vid, str(tmp_path / "a"))
    back = video_io.read_video_sequence(str(tmp_path / "a"))
    assert back.shape == vid.shape and np.array_equal(back, vid)
    video_io.save_video_sequence(vid, str(tmp_path / "b"), fmt="%03d.npy")
    assert np.array_equal(video_io.read_video_sequence(str(tmp_path / "b")), vid)
    rep = video_io.compare_report(vid + 1, vid, clean=vid)
    assert rep["Ave Rel. Error"] > 0 and np.isinf(rep["other_psnr"]) and abs(rep["our_psnr"] - 48.13) < 0.01
    with pytest.raises(FileNotFoundError):
        video_io.read_video_sequence(str(tmp_path / "none"))


def test_graph_row_buckets_hold_the_draw_and_are_few():
    """schedule._graph_rows: the row count of a captured round is one of ~10 buckets per step (cap, 3/4 cap, 1/2 cap, ...,
    >= 512), never smaller than the rows the draw needs."""
    from vnlb_b200.schedule import _graph_rows
    cap = 16384
    seen = set()
    for need in list(range(1, 2000, 37)) + list(range(2000, cap + 1, 311)) + [cap]:
        rows = _graph_rows(need, cap)
        assert rows >= min(need, cap) and rows <= cap and rows >= 512
        seen.add(rows)
        if rows > 512 and rows * 3 // 4 >= 512:
            assert need > rows * 3 // 4 or rows == 512 or need > rows // 2     # the next smaller bucket would not hold it
    assert len(seen) <= 11
