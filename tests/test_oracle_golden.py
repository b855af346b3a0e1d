"""Pin the oracle against outputs of the reference's own NumPy functions (tests/golden)."""
import numpy as np
import pytest

from oracle import lcn_oracle as O


@pytest.mark.parametrize("knn", [1, 2, 3, 4, 5])
def test_neighbour_matrix_bit_exact(golden, knn):
    ref = golden[f"neighbour_knn{knn}"]
    got = O.get_neighbour_matrix_by_hand(knn=knn)
    assert got.dtype == ref.dtype == np.float32
    assert got.tobytes() == ref.tobytes()
    assert int(got.sum()) == {1: 57, 2: 111, 3: 175, 4: 237, 5: 273}[knn]


def test_exponential_matrix_bit_exact(golden):
    ref = golden["exponential"]
    got = O.get_exponential_matrix()
    assert got.dtype == np.float32 and got.tobytes() == ref.tobytes()
    vals, counts = np.unique(got, return_counts=True)
    assert dict(zip(vals.tolist(), counts.tolist())) == {1.0: 17, .5: 36, .25: 50, .125: 60, .0625: 66,
                                                         .03125: 40, .015625: 20}


def test_procrustes_known_answer(golden):
    gt = np.arange(51, dtype=np.float64).reshape(17, 3) ** 1.5
    pred = gt[:, ::-1] * 0.9 + np.sin(np.arange(51)).reshape(17, 3) * 7
    d, Z, tf = O.procrustes(gt, pred)
    assert abs(d - 0.002549649086) < 1e-11
    assert abs(tf["scale"] - 1.113359602576) < 1e-11
    assert abs(np.sqrt(((Z - gt) ** 2).sum(1)).mean() - 9.424886756111) < 1e-10
    np.testing.assert_allclose(Z, golden["ka_proc_Z"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(tf["rotation"], golden["ka_proc_R"], rtol=0, atol=1e-12)


def test_image_to_camera_frame_known_answer(golden):
    got = O.image_to_camera_frame(np.arange(51, dtype=np.float64).reshape(17, 3) * 3 + 100,
                                  box=[200, 150, 800, 750], camera={"cx": 512, "cy": 515, "fx": 1145, "fy": 1144},
                                  rootIdx=0, root_depth=5000)
    np.testing.assert_allclose(got[0], [-1926.05337538, -1927.73698847, 5352.74542429], atol=1e-7)
    np.testing.assert_allclose(got[16], [-1365.03207899, -1366.22528885, 5831.94675541], atol=1e-7)
    assert np.array_equal(got, golden["ka_i2c"])


def test_eval_loop_matches_reference(golden):
    g = golden
    n = len(g["ev_pred"])
    # rows with i % 16 == 5 had their camera-frame pose replaced by a reflected gt in make_golden.py
    normal = np.array([i % 16 != 5 for i in range(n)])
    p1 = O.eval_errors(g["ev_pred"], g["ev_gt"], g["ev_box"], g["ev_cam"], g["ev_root_depth"], False)
    np.testing.assert_allclose(p1[normal], g["ev_err_p1"][normal], rtol=0, atol=1e-9)
    p2 = O.eval_errors(g["ev_pred"], g["ev_gt"], g["ev_box"], g["ev_cam"], g["ev_root_depth"], True)
    np.testing.assert_allclose(p2[normal], g["ev_err_p2"][normal], rtol=0, atol=1e-8)
    # reflected cases: procrustes accepts det(R) = -1 and aligns (almost) exactly
    for i in np.nonzero(~normal)[0]:
        al = O.align_to_gt(g["ev_camframe"][i], g["ev_gt"][i])
        np.testing.assert_allclose(al, g["ev_aligned"][i], rtol=0, atol=1e-8)
        assert np.sqrt(((al - g["ev_gt"][i]) ** 2).sum(1)).max() < 1e-6
        assert np.linalg.det(O.procrustes(g["ev_gt"][i], g["ev_camframe"][i])[2]["rotation"]) < 0


def test_augmentations_match_reference(golden):
    g = golden
    assert np.array_equal(O.flip_data(g["aug_x2"]), g["aug_flip2"])
    assert np.array_equal(O.flip_data(g["aug_x3"]), g["aug_flip3"])
    np.testing.assert_allclose(O.rotate_data(g["aug_x2"], 37.0), g["aug_rot2"], atol=1e-14)
    np.testing.assert_allclose(O.rotate_data(g["aug_x3"], 37.0), g["aug_rot3"], atol=1e-14)
    np.testing.assert_allclose(O.translation_data(g["aug_x2"], 0.07), g["aug_tr2"], atol=1e-15)
    np.testing.assert_allclose(O.translation_data(g["aug_x3"], 0.07), g["aug_tr3"], atol=1e-15)


def test_mask_values_at_init():
    """in-support entry of column j equals e/(k_j*e + 17 - k_j)."""
    cfg = O.LcnConfig(neighbour_matrix=O.get_neighbour_matrix_by_hand(knn=3))
    p = O.init_params(cfg, seed=1)
    m = O.mask_values(cfg, p)
    sup = cfg.support()
    for j in range(17):
        k = int(sup[:, j].sum())
        np.testing.assert_allclose(m[sup[:, j], j], np.e / (k * np.e + 17 - k), rtol=1e-14)
        assert np.all(m[~sup[:, j], j] == 0)
    cfg_bad = O.LcnConfig(init_type="ones")
    with pytest.raises(ValueError):
        O.init_params(cfg_bad)


def test_predict_zero_pads_last_batch():
    cfg = O.LcnConfig(F=8, num_layers=1, neighbour_matrix=O.get_neighbour_matrix_by_hand(knn=2))
    p = O.init_params(cfg, seed=3)
    rng = np.random.default_rng(0)
    x = rng.normal(0, 0.3, (11, 34))
    pr = O.predict(cfg, p, x, batch_size=4)
    pad = np.zeros((4, 34)); pad[:3] = x[8:]
    np.testing.assert_allclose(pr[8:], O.forward(cfg, p, pad)[0][:3], atol=1e-15)
    np.testing.assert_allclose(pr[:4], O.forward(cfg, p, x[:4])[0], atol=1e-15)


def test_tta_undo_matches_reference(golden):
    g = golden
    got = O.undo(g["aug_undo_in"], {"f": 1, "r": 2, "t": 3}, number_actions=3, angle=180, translation=0.07)
    np.testing.assert_allclose(got, g["aug_undo_out"], atol=1e-14)
    got = O.undo(g["aug_undo_in"][:18], {"r": 1, "f": 2}, number_actions=2, angle=180, translation=0.0)
    np.testing.assert_allclose(got, g["aug_undo_fr_out"], atol=1e-14)


@pytest.mark.parametrize("tag,kw", [("best", {}), ("noscale", {"scaling": False}), ("norefl", {"reflection": False}),
                                    ("refl", {"reflection": True}),
                                    ("noscale_norefl", {"scaling": False, "reflection": False})])
def test_procrustes_every_option_matches_reference(golden, tag, kw):
    """tools.procrustes(A, B, scaling, reflection) (tools/tools.py:96-181): the oracle against the reference's own output
    for every option of the signature, including a reflected input where the forced variants differ from 'best'."""
    A, B = golden["proc_A"], golden["proc_B"]
    for i in range(len(A)):
        d, Z, tf = O.procrustes(A[i], B[i], **kw)
        assert abs(d - golden[f"proc_{tag}_d"][i]) < 1e-10
        np.testing.assert_allclose(Z, golden[f"proc_{tag}_Z"][i], rtol=0, atol=1e-8)
        np.testing.assert_allclose(tf["rotation"], golden[f"proc_{tag}_R"][i], rtol=0, atol=1e-10)
        assert abs(tf["scale"] - golden[f"proc_{tag}_scale"][i]) < 1e-10
        np.testing.assert_allclose(tf["translation"], golden[f"proc_{tag}_t"][i], rtol=0, atol=1e-7)
    assert not np.allclose(golden["proc_best_Z"][3], golden["proc_norefl_Z"][3])     # the options do something


def test_datareader_normalise_denormalise_pinned_to_reference(golden):
    """DataReader.read_2d / read_3d / denormalize (tools/data.py:338-489), run by make_golden.py on a synthetic dataitem
    list covering the three camera-resolution classes."""
    for split in ("train", "test"):
        cams = golden[f"dr_{split}_cam"]
        res = np.array([O.camera_resolution(c) for c in cams], dtype=np.float64)
        j3d = golden[f"dr_{split}_j3d"]
        np.testing.assert_allclose(O.normalize_2d(j3d, res[:, 0], res[:, 1]), golden[f"dr_x_{split}"], rtol=0, atol=1e-13)
        np.testing.assert_allclose(O.normalize_3d(j3d, res[:, 0], res[:, 1]), golden[f"dr_y_{split}"], rtol=0, atol=1e-13)
    res = np.array([O.camera_resolution(c) for c in golden["dr_test_cam"]], dtype=np.float64)
    den = O.denormalize(golden["dr_y_test"], res[:, 0], res[:, 1])
    np.testing.assert_allclose(den, golden["dr_denorm"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(den, golden["dr_test_j3d"], rtol=0, atol=1e-9)     # denormalize inverts read_3d


@pytest.mark.parametrize("k", [3, 2])
def test_rotate_180_known_answer_of_the_reference_tests(k):
    """The two cases of the reference's own tools/tests.py (TestRotateData): a 180-degree rotation about joint 0 maps
    xy -> -(xy - pivot) + pivot and leaves the third coordinate alone.  (Those tests also expect the batch to double,
    which tools/data.py:289-322 no longer does; the arithmetic they pin is what is checked here.)"""
    data = np.zeros((1, 17, k))
    for i in range(17):
        data[0, i] = [i, i + 1, 1][:k]
    rot = O.rotate_data(data, angle=180)
    assert rot.shape == data.shape
    pivot = data[0, 0, :2]
    np.testing.assert_allclose(rot[0, :, :2], -(data[0, :, :2] - pivot) + pivot, atol=1e-6)
    if k == 3:
        np.testing.assert_allclose(rot[0, :, 2], data[0, :, 2])
