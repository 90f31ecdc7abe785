"""Dataloader -> device staging.

CPU: the numpy oracle against the golden vectors produced by the reference's own BaseDataset.__getitem__
(tests/golden/dataloader_ref.npz), the PFM reader, the raw dataset's index logic.  GPU: the staging kernels through the
C ABI against the golden vectors and the oracle, plus size-independent properties at DAVIS-sized frames."""
import os
import sys
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dataloader as OD  # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "dataloader_ref.npz"))
KEYS = ("img", "imgn", "mask", "flow", "flown", "occ", "occn", "cam", "camn", "pps")
S = int(GOLD["img_size"])
VIDEOS = ("", "b_")      # object inside the frame / at the frame corner (the crop reaches outside: constant border)


def write_pfm(path, data):
    data = np.asarray(data, np.float32)
    with open(path, "wb") as fh:
        fh.write(b"PF\n" if data.ndim == 3 else b"Pf\n")
        fh.write(f"{data.shape[1]} {data.shape[0]}\n".encode())
        fh.write(b"-1.0\n")
        np.flipud(data).astype("<f4").tofile(fh)


# ------------------------------------------------------------------------------------------------ CPU
@pytest.mark.parametrize("pre", VIDEOS)
def test_oracle_reproduces_the_reference_dataloader(pre):
    if pre == "b_":
        assert (GOLD["b_item0/pps"] < 0).any()          # the crop does leave the frame in this video
    for it in GOLD[pre + "items"]:
        r = OD.preprocess_pair(*OD.pair_inputs(GOLD, int(it), pre), S)
        for k in KEYS:
            ref = GOLD[f"{pre}item{it}/{k}"]
            assert r[k].shape == ref.shape, (it, k)
            if k in ("img", "imgn"):          # float64 image path: summation order of the masked mean
                np.testing.assert_allclose(r[k], ref, rtol=0, atol=1e-15, err_msg=f"{pre}{it}/{k}")
            else:                             # float32 flow / occlusion, masks, crop boxes: bit-exact
                assert np.array_equal(r[k], ref), (pre, it, k)


def test_oracle_resize_taps_are_cv2s():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(0)
    for src, dst in ((37, 64), (128, 64), (200, 64), (64, 64), (90, 256)):
        a = rng.uniform(size=(src, src, 3)).astype(np.float32)
        np.testing.assert_allclose(OD._resize_linear(a, dst), cv2.resize(a, (dst, dst), interpolation=cv2.INTER_LINEAR),
                                   rtol=0, atol=2e-7)
        m = (rng.uniform(size=(src, src)) > 0.5).astype(np.uint8)
        assert np.array_equal(OD._resize_nearest(m, dst), cv2.resize(m, (dst, dst), interpolation=cv2.INTER_NEAREST))


def test_pfm_reader_round_trip(tmp_path):
    from lasr_b200.dataloader import readPFM
    rng = np.random.RandomState(1)
    for shape in ((5, 7, 3), (4, 9)):
        a = rng.normal(size=shape).astype(np.float32)
        p = str(tmp_path / "a.pfm"); write_pfm(p, a)
        d, scale = readPFM(p)
        assert scale == 1.0 and np.array_equal(d, a)
        with open(p, "wb") as fh:           # big-endian file
            fh.write(b"PF\n" if a.ndim == 3 else b"Pf\n"); fh.write(f"{a.shape[1]} {a.shape[0]}\n".encode()); fh.write(b"2.0\n")
            np.flipud(a).astype(">f4").tofile(fh)
        d, scale = readPFM(p)
        assert scale == 2.0 and np.array_equal(d, a)
    (tmp_path / "bad.pfm").write_bytes(b"P6\n1 1\n1\n")
    with pytest.raises(Exception):
        readPFM(str(tmp_path / "bad.pfm"))


def test_raw_dataset_decodes_what_the_reference_decodes(tmp_path):
    cv2 = pytest.importorskip("cv2")
    from lasr_b200.dataloader import BaseDataset, collate_raw
    n = len(GOLD["frames_bgr"])
    L = dict(img=[], mask=[], fw=[], bw=[])
    for t in range(n):
        p = {k: str(tmp_path / f"{k}-{t}.{'png' if k in ('img', 'mask') else 'pfm'}") for k in L}
        p["fw"], p["bw"] = str(tmp_path / f"flo-fw-{t}.pfm"), str(tmp_path / f"flo-bw-{t}.pfm")
        cv2.imwrite(p["img"], GOLD["frames_bgr"][t]); cv2.imwrite(p["mask"], GOLD["masks"][t])
        for d in ("fw", "bw"):
            write_pfm(p[d], GOLD[f"flow_{d}"][t]); write_pfm(p[d].replace("flo-", "occ-"), GOLD[f"occ_{d}"][t])
        for k in L:
            L[k].append(p[k])
    ds = BaseDataset(types.SimpleNamespace(img_size=S))
    ds.imglist, ds.masklist, ds.flowfwlist, ds.flowbwlist = L["img"], L["mask"], L["fw"], L["bw"]
    ds.baselist, ds.directlist = list(GOLD["baselist"]), list(GOLD["directlist"])
    ds.dframe, ds.can_frame, ds.num_imgs = 1, 0, len(GOLD["baselist"])
    assert len(ds) == len(GOLD["baselist"])
    for it in GOLD["items"]:
        raw = ds[int(it)]
        want = OD.pair_inputs(GOLD, int(it))
        assert np.array_equal(raw["frames_bgr"][0], want[0]) and np.array_equal(raw["frames_bgr"][1], want[4])
        assert np.array_equal(raw["masks"][0], want[1]) and np.array_equal(raw["masks"][1], want[5])
        assert np.array_equal(raw["flows"][0], want[2]) and np.array_equal(raw["flows"][1], want[6])
        assert np.array_equal(raw["occs"][0], want[3]) and np.array_equal(raw["occs"][1], want[7])
        assert raw["id0"] == int(GOLD[f"item{it}/id0"]) and raw["id1"] == int(GOLD[f"item{it}/id1"])
    batch = collate_raw([ds[int(i)] for i in GOLD["items"]])
    assert tuple(batch["frames_bgr"].shape) == (3, 2) + GOLD["frames_bgr"].shape[1:] and batch["flows"].dtype == torch.float32


def test_masks_at_another_resolution_are_resized_and_eroded_like_the_reference(tmp_path):
    """Video "m_" of the golden file stores its masks at half resolution: the reference resizes them to the frame and
    erodes them (two iterations for frame 0, one for its partner, dataloader/vidbase.py:67-76) before everything else.
    The raw dataset must hand over masks from which the oracle reproduces the reference's items."""
    cv2 = pytest.importorskip("cv2")
    from lasr_b200.dataloader import BaseDataset
    pre = "m_"
    n = len(GOLD[pre + "frames_bgr"])
    L = dict(img=[], mask=[], fw=[], bw=[])
    for t in range(n):
        p = dict(img=str(tmp_path / f"img-{t}.png"), mask=str(tmp_path / f"mask-{t}.png"),
                 fw=str(tmp_path / f"flo-fw-{t}.pfm"), bw=str(tmp_path / f"flo-bw-{t}.pfm"))
        cv2.imwrite(p["img"], GOLD[pre + "frames_bgr"][t]); cv2.imwrite(p["mask"], GOLD[pre + "masks"][t])
        for d in ("fw", "bw"):
            write_pfm(p[d], GOLD[pre + f"flow_{d}"][t]); write_pfm(p[d].replace("flo-", "occ-"), GOLD[pre + f"occ_{d}"][t])
        for k in L:
            L[k].append(p[k])
    assert GOLD[pre + "masks"].shape[1] * 2 == GOLD[pre + "frames_bgr"].shape[1]
    camlist = []
    for t in range(n):                   # this video also has per-frame camera files (vidbase.py:174-181)
        camlist.append(str(tmp_path / f"cam-{t}.txt"))
        np.savetxt(camlist[-1], GOLD[pre + "cams"][t])
    ds = BaseDataset(types.SimpleNamespace(img_size=S))
    ds.camlist = camlist
    ds.imglist, ds.masklist, ds.flowfwlist, ds.flowbwlist = L["img"], L["mask"], L["fw"], L["bw"]
    ds.baselist, ds.directlist = list(GOLD[pre + "baselist"]), list(GOLD[pre + "directlist"])
    ds.dframe, ds.can_frame, ds.num_imgs = 1, 0, len(GOLD[pre + "baselist"])
    for it in GOLD[pre + "items"]:
        raw = ds[int(it)]
        assert raw["masks"].shape[1:] == raw["frames_bgr"].shape[1:3] and raw["masks"].dtype == np.uint8
        r = OD.preprocess_pair(raw["frames_bgr"][0], raw["masks"][0], raw["flows"][0], raw["occs"][0],
                               raw["frames_bgr"][1], raw["masks"][1], raw["flows"][1], raw["occs"][1], S)
        for k in KEYS:
            ref = GOLD[f"{pre}item{it}/{k}"]
            if k in ("img", "imgn"):
                np.testing.assert_allclose(r[k], ref, rtol=0, atol=1e-15, err_msg=f"{it}/{k}")
            elif k in ("cam", "camn"):      # entries 1..6 come from the camera file, entry 0 is the crop's focal scale
                slot = 0 if k == "cam" else 1
                assert r[k][0] == ref[0]
                np.testing.assert_allclose(raw["cam_init"][slot][1:], ref[1:], rtol=1e-6)
            else:
                assert np.array_equal(r[k], ref), (it, k)
        np.testing.assert_allclose(raw["depth"], [float(GOLD[f"{pre}item{it}/depth"][0]), float(GOLD[f"{pre}item{it}/depthn"][0])])


def test_cameras_default_to_identity_without_camera_files():
    from lasr_b200.dataloader import collate_raw
    item = {"frames_bgr": np.zeros((2, 4, 4, 3), np.uint8), "masks": np.zeros((2, 4, 4), np.uint8),
            "flows": np.zeros((2, 4, 4, 3), np.float32), "occs": np.zeros((2, 4, 4), np.float32), "shape": np.asarray([4, 4]),
            "cam_init": np.tile(np.asarray([1., 0, 0, 1, 0, 0, 0], np.float32), (2, 1)), "depth": np.zeros(2, np.float32),
            "inds": 0, "indsn": 0, "id0": 0, "id1": 1, "is_canonical": True, "is_canonicaln": False, "dataid": 0}
    b = collate_raw([item, item])
    assert tuple(b["cam_init"].shape) == (2, 2, 7) and tuple(b["depth"].shape) == (2, 2)


def test_pair_selection_follows_the_reference():
    from lasr_b200.dataloader import pair_frame_indices
    for pre in VIDEOS:
        for it in GOLD[pre + "items"]:
            i0, i1, fwd = pair_frame_indices(GOLD[pre + "baselist"][it], GOLD[pre + "directlist"][it], 1)
            assert i0 == int(GOLD[f"{pre}item{it}/id0"]) and i1 == int(GOLD[f"{pre}item{it}/id1"])
            assert fwd == (int(GOLD[pre + "directlist"][it]) == 1)


def _raw_batch(items, pre=""):
    ins = [OD.pair_inputs(GOLD, int(i), pre) for i in items]
    st = lambda a, b: np.stack([np.stack([x[a], x[b]]) for x in ins])
    return st(0, 4), st(1, 5), st(2, 6).astype(np.float32), st(3, 7).astype(np.float32)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("pre", VIDEOS)
def test_staging_kernels_match_the_reference_dataloader(pre):
    from lasr_b200.dataloader import stage_pairs
    dev = torch.device("cuda", 0)
    items = [int(i) for i in GOLD[pre + "items"]]
    fr, mk, fl, oc = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in _raw_batch(items, pre))
    out = stage_pairs(fr, mk, fl, oc, S)
    torch.cuda.synchronize()
    for p, it in enumerate(items):
        for k in KEYS:
            got = out[k][p].cpu().numpy().astype(np.float64)
            ref = GOLD[f"{pre}item{it}/{k}"].astype(np.float64)
            assert got.shape == ref.shape, (it, k)
            if k in ("mask", "pps", "cam", "camn"):
                np.testing.assert_allclose(got, ref, rtol=1e-7, atol=0, err_msg=f"{it}/{k}")      # exact up to f32 storage
            elif k in ("flow", "flown"):
                assert np.array_equal(got[2], ref[2]), (it, k)                                      # validity channel
                np.testing.assert_allclose(got[:2], ref[:2], rtol=1e-5, atol=2e-6, err_msg=f"{it}/{k}")
            else:                                                                                   # float32 pipeline vs float64 / float32
                np.testing.assert_allclose(got, ref, rtol=1e-5, atol=2e-6, err_msg=f"{it}/{k}")
        b = out["boxes"][p].cpu().numpy()
        for slot, m in enumerate((OD.pair_inputs(GOLD, it, pre)[1], OD.pair_inputs(GOLD, it, pre)[5])):
            assert tuple(b[slot][:3]) == OD.crop_box(m) and b[slot][3] == int((m > 0).sum())


@pytest.mark.gpu
def test_staging_properties_at_davis_size_and_through_the_stager():
    from lasr_b200.dataloader import DeviceStager
    from lasr_b200.dataloader.staging import check_boxes, stage_pairs
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(0)
    P, H, W, S2 = 4, 480, 854, 256
    fr = torch.randint(0, 256, (P, 2, H, W, 3), generator=g, dtype=torch.uint8)
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    mk = torch.zeros(P, 2, H, W, dtype=torch.uint8)
    for p in range(P):
        for s in range(2):      # masks touching the frame border on some items: the crop reaches outside
            cx, cy = 120 + 200 * p + 15 * s, 100 + 90 * p
            mk[p, s] = ((((xx - cx) / 110.0) ** 2 + ((yy - cy) / 95.0) ** 2) < 1).to(torch.uint8) * 255
    fl = torch.randn(P, 2, H, W, 3, generator=g); fl[..., 2] = (torch.rand(P, 2, H, W, generator=g) > 0.1).float()
    oc = torch.rand(P, 2, H, W, generator=g) * 20
    st = DeviceStager(P, H, W, S2)
    out = st.submit({"frames_bgr": fr, "masks": mk, "flows": fl, "occs": oc}).result()
    torch.cuda.synchronize()
    check_boxes(out["boxes"])
    assert st.h2d_bytes == fr.numel() + mk.numel() + 4 * (fl.numel() + oc.numel())
    # same result as the functional form on resident tensors; deterministic
    ref = stage_pairs(fr.to(dev), mk.to(dev), fl.to(dev), oc.to(dev), S2)
    for k in KEYS:
        assert torch.equal(out[k], ref[k]), k
    # properties: image values stay inside the convex hull of [0,1]; mask and validity are binary; the crop is square and
    # centred on the mask box; pp = centre - half size; focal = S / (2L)
    assert float(out["img"].min()) >= 0 and float(out["img"].max()) <= 1
    assert set(torch.unique(out["mask"]).tolist()) <= {0.0, 1.0} and set(torch.unique(out["flow"][:, 2]).tolist()) <= {0.0, 1.0}
    b = out["boxes"].cpu()
    for p in range(P):
        for s in range(2):
            cx, cy, L, cnt = (int(v) for v in b[p, s])
            assert OD.crop_box(mk[p, s].numpy()) == (cx, cy, L) and cnt == int((mk[p, s] > 0).sum())
            assert torch.allclose(out["pps"][p, s].cpu(), torch.tensor([float(cx - L), float(cy - L)]))
            cam = (out["cam"] if s == 0 else out["camn"])[p].cpu()
            assert abs(float(cam[0]) - S2 / (2.0 * L)) < 1e-6 and float(cam[3]) == 1.0
    # oracle on one pair at this size
    r = OD.preprocess_pair(fr[1, 0].numpy(), mk[1, 0].numpy(), fl[1, 0].numpy(), oc[1, 0].numpy(),
                           fr[1, 1].numpy(), mk[1, 1].numpy(), fl[1, 1].numpy(), oc[1, 1].numpy(), S2)
    for k in ("img", "imgn", "flow", "flown", "occ", "occn"):
        np.testing.assert_allclose(out[k][1].cpu().numpy(), r[k], rtol=1e-5, atol=3e-6, err_msg=k)
    assert np.array_equal(out["mask"][1].cpu().numpy(), r["mask"])
    # an empty mask is reported, not silently staged
    mk2 = mk.clone(); mk2[2, 1] = 0
    with pytest.raises(ValueError):
        check_boxes(stage_pairs(fr.to(dev), mk2.to(dev), fl.to(dev), oc.to(dev), S2)["boxes"])
    with pytest.raises((TypeError, RuntimeError)):
        stage_pairs(fr, mk, fl, oc, S2)            # CPU tensors are refused loudly


@pytest.mark.gpu
@pytest.mark.parametrize("pre", VIDEOS)
def test_resident_video_stages_once_and_serves_batches(pre):
    """The whole video in HBM, every (frame, direction) item staged on first use, batches by index_select: the items the
    reference's dataset returns, in any order and repeatedly."""
    from lasr_b200.dataloader import ResidentVideo, stage_pairs
    G = lambda k: GOLD[pre + k]
    rv = ResidentVideo(G("frames_bgr"), G("masks"), G("flow_fw"), G("occ_fw"), G("flow_bw"), G("occ_bw"), S, chunk=3)
    items = [int(i) for i in G("items")]
    base = [int(G("baselist")[i]) for i in items]; dirs = [int(G("directlist")[i]) for i in items]
    out = rv.stage(base, dirs)
    for p, it in enumerate(items):
        for k in KEYS:
            got = out[k][p].cpu().numpy().astype(np.float64); ref = GOLD[f"{pre}item{it}/{k}"].astype(np.float64)
            np.testing.assert_allclose(got, ref, rtol=1e-5, atol=2e-6, err_msg=f"{pre}{it}/{k}")
    again = rv.stage(base[::-1], dirs[::-1])                       # served from the staged items, any order
    for k in KEYS:
        assert torch.equal(again[k], out[k].flip(0)), k
    dev = torch.device("cuda", 0)
    direct = stage_pairs(*(torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in _raw_batch(items, pre)), S)
    for k in KEYS:
        assert torch.equal(out[k], direct[k]), k                   # identical to staging the same pairs directly
    with pytest.raises(IndexError):
        rv.stage([0], [0])                                         # frame 0 has no predecessor
