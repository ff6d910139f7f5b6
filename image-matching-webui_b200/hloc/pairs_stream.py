"""Batch exporters of the pair stream -- the GPU-batched counterparts of the reference's file-to-file drivers
`extract_features.main` (imcui/hloc/extract_features.py:173-248) and `match_features.main` / `match_from_paths` /
`find_unique_new_pairs` / `writer_fn` (match_features.py:73-186): same inputs (image folder or list, `pairs.txt`), same
group / dataset names and dtypes in the output files (utils/store.py), same skip / overwrite / de-duplication rules.

What changes is the schedule.  The reference walks one image (one pair) at a time through a DataLoader, syncs on `.cpu()`
and converts on the host.  Here
  * images are decoded by a small thread pool, same-size frames are grouped, and each group goes through ONE upload, ONE
    imw_preprocess pass and ONE extractor forward; keypoints are rescaled to the original frame and everything is cast to
    fp16 on the device, so one D2H per group feeds the writer;
  * pairs are packed `batch` at a time into padded [2P, cap, .] buffers by a loader thread (features read as stored,
    fp16 -> fp32 as FeaturePairsDataset does), matched with one `match_batch` call per batch (hloc/matchers/*), narrowed to
    int16 / fp16 on the device and written by a writer thread while the next batch is in flight.
"""
import queue
import threading
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import cv2
import numpy as np
import torch

from . import extractors, logger, matchers
from .. import ops
from .utils.base_model import dynamic_load
from .utils.parsers import names_to_pair, names_to_pair_old, parse_image_lists, parse_retrieval
from .utils.store import open_store

IMAGE_GLOBS = ["*.jpg", "*.png", "*.jpeg", "*.JPG", "*.PNG"]     # extract_features.py:45


# =========================================================================================================
# features
# =========================================================================================================
def list_images(image_dir, image_list=None):
    """extract_features.py:52-75: explicit list (file, glob of files, or iterable of names) or every image under the root"""
    root = Path(image_dir)
    if image_list is None:
        paths = sorted({p for g in IMAGE_GLOBS for p in root.glob("**/" + g)})
        if not paths:
            raise ValueError(f"Could not find any image in root: {root}.")
        return [p.relative_to(root).as_posix() for p in paths]
    if isinstance(image_list, (str, Path)):
        names = parse_image_lists(Path(image_list))
    else:
        names = [p.as_posix() if isinstance(p, Path) else p for p in image_list]
    for n in names:
        if not (root / n).exists():
            raise ValueError(f"Image {n} does not exists in root: {root}.")
    return names


def read_image(path, grayscale=False):
    """hloc/utils/io.py:11-21: decoded by OpenCV, gray at decode time when the conf asks for it, BGR -> RGB otherwise"""
    image = cv2.imread(str(path), cv2.IMREAD_GRAYSCALE if grayscale else cv2.IMREAD_COLOR)
    if image is None:
        raise ValueError(f"Cannot read image {path}.")
    return image if grayscale or image.ndim != 3 else np.ascontiguousarray(image[:, :, ::-1])


def _dataset_size(shape_hw, pre):
    """ImageDataset.__getitem__ (extract_features.py:82-87): resize so that the longer side is resize_max when the image is
    larger (or always, with force_resize).  -> (w, h) or None"""
    h, w = shape_hw
    rm = pre.get("resize_max")
    if rm and (pre.get("force_resize") or max(w, h) > rm):
        scale = rm / max(w, h)
        return tuple(int(round(x * scale)) for x in (w, h))
    return None


@torch.no_grad()
def export_features(conf, image_dir, export_dir=None, as_half=True, image_list=None, feature_path=None, overwrite=False,
                    batch=32, device="cuda", decode_threads=8):
    """extract_features.main (:173-248).  conf: an entry of confs_dict["extractors"].  Returns the feature file path."""
    names = list_images(image_dir, image_list)
    if feature_path is None:
        feature_path = Path(export_dir, conf["output"] + ".h5")
    feature_path = Path(feature_path)
    feature_path.parent.mkdir(exist_ok=True, parents=True)
    if feature_path.exists() and not overwrite:
        with open_store(feature_path, "r") as st:
            done = set(st.groups())
        names = [n for n in names if n not in done]
    if not names:
        logger.info("Skipping the extraction.")
        return feature_path
    pre = {"grayscale": False, "resize_max": None, "force_resize": False, **conf["preprocessing"]}
    model = dynamic_load(extractors, conf["model"]["name"])(conf["model"]).eval().to(device)
    noise = getattr(model, "detection_noise", 1)
    root = Path(image_dir)
    store = open_store(feature_path, "w" if overwrite else "a")
    pool = ThreadPoolExecutor(decode_threads)
    try:
        for i0 in range(0, len(names), batch):
            chunk = names[i0:i0 + batch]
            frames = list(pool.map(lambda n: read_image(root / n, pre["grayscale"]), chunk))
            groups = {}
            for j, f in enumerate(frames):
                groups.setdefault(f.shape, []).append(j)
            for shape, idx in groups.items():
                target = _dataset_size(shape[:2], pre)
                pconf = {"grayscale": pre["grayscale"], "resize_max": 0, "dfactor": 1, "force_resize": target is not None,
                         "width": target[0] if target else 0, "height": target[1] if target else 0}
                dev_u8 = torch.from_numpy(np.stack([frames[j] for j in idx])).to(device)
                image = ops.preprocess(dev_u8, pconf)                                  # [n,C,H',W'] in [0,1]
                pred = model({"image": image})
                orig = np.array(shape[:2][::-1])
                size = np.array([image.shape[3], image.shape[2]])
                scales = (orig / size).astype(np.float32)                              # extract_features.py:212-214
                sc_t = torch.from_numpy(scales).to(device)[None]
                # device side: rescale / cast every output of the group, then ONE packed D2H per dtype for the whole group
                entries, attrs = [], {"keypoints": {"uncertainty": noise * scales.mean()}} if "keypoints" in pred else {}
                for b, j in enumerate(idx):
                    for k, v in pred.items():
                        t = v[b]
                        if k == "keypoints" and t.shape[0] > 0:
                            t = ops.rescale_keypoints(t.float()[None].contiguous(), sc_t)[0]
                        elif k == "scales":
                            t = t * float(scales.mean())
                        if as_half and t.dtype == torch.float32:
                            t = t.half()                                               # round-to-nearest, as ndarray.astype(float16)
                        entries.append((j, k, t.contiguous()))
                host = {}
                for dt in {t.dtype for _, _, t in entries}:
                    sel = [(j, k, t) for j, k, t in entries if t.dtype == dt]
                    flat = torch.cat([t.reshape(-1) for _, _, t in sel]).cpu().numpy()
                    o = 0
                    for j, k, t in sel:
                        host.setdefault(j, {})[k] = flat[o:o + t.numel()].reshape(tuple(t.shape)).copy()
                        o += t.numel()
                for j in idx:
                    store.write_group(chunk[j], {"image_size": orig, **host.get(j, {})}, attrs)
    finally:
        pool.shutdown()
        store.close()
    logger.info("Finished exporting features.")
    return feature_path


# =========================================================================================================
# matches
# =========================================================================================================
def find_unique_new_pairs(pairs_all, match_path=None):
    """match_features.py:117-138: drop (j, i) when (i, j) is present, and pairs already stored under either key style.
    (The reference builds a set, so its order is arbitrary; first-seen order is kept here.)"""
    seen, pairs = set(), []
    for i, j in pairs_all:
        if (j, i) not in seen and (i, j) not in seen:
            seen.add((i, j))
            pairs.append((i, j))
    if match_path is not None and Path(match_path).exists():
        with open_store(match_path, "r") as st:
            pairs = [(i, j) for i, j in pairs
                     if not any(k in st for k in (names_to_pair(i, j), names_to_pair(j, i), names_to_pair_old(i, j), names_to_pair_old(j, i)))]
    return pairs


class _PairLoader(threading.Thread):
    """Reads stored features and packs `batch` pairs into pinned, padded buffers (FeaturePairsDataset :47-70 for many pairs)."""

    def __init__(self, pairs, path_q, path_r, batch, out_q):
        super().__init__(daemon=True)
        self.pairs, self.path_q, self.path_r, self.batch, self.q = pairs, path_q, path_r, batch, out_q
        self.error = None

    def run(self):
        try:
            sq = open_store(self.path_q, "r")
            sr = sq if Path(self.path_r) == Path(self.path_q) else open_store(self.path_r, "r")
            for i0 in range(0, len(self.pairs), self.batch):
                chunk = self.pairs[i0:i0 + self.batch]
                feats = [[{k: st.read(n, k) for k in st.datasets(n)} for st, n in ((sq, a), (sr, b))] for a, b in chunk]
                nmax = max(max(len(f["keypoints"]) for f in pr) for pr in feats)
                cap = max(128, (nmax + 127) // 128 * 128)
                dim = feats[0][0]["descriptors"].shape[0]
                S = 2 * len(chunk)
                kp = torch.zeros(S, cap, 2).pin_memory(); ds = torch.zeros(S, cap, dim).pin_memory(); sc = torch.zeros(S, cap).pin_memory()
                counts = torch.zeros(S, dtype=torch.int32); wh = torch.zeros(S, 2, dtype=torch.int32)
                extra = {k: torch.zeros(S, cap).pin_memory() for k in ("scales", "oris") if k in feats[0][0]}   # add_scale_ori features
                for p, pr in enumerate(feats):
                    for side, f in enumerate(pr):
                        n = len(f["keypoints"])
                        z = 2 * p + side
                        kp[z, :n] = torch.from_numpy(f["keypoints"].astype(np.float32))
                        ds[z, :n] = torch.from_numpy(np.ascontiguousarray(f["descriptors"].T).astype(np.float32))
                        if "scores" in f:
                            sc[z, :n] = torch.from_numpy(f["scores"].astype(np.float32))
                        for k, t in extra.items():
                            t[z, :n] = torch.from_numpy(f[k].astype(np.float32))
                        counts[z] = n
                        wh[z] = torch.from_numpy(np.asarray(f["image_size"]).astype(np.int32))    # "some matchers ... only use its size"
                self.q.put((chunk, {"keypoints": kp, "descriptors": ds, "scores": sc, "counts": counts, "image_wh": wh, **extra}))
        except Exception as e:  # noqa: BLE001  (re-raised in the consumer)
            self.error = e
        finally:
            self.q.put(None)


@torch.no_grad()
def match_from_paths(conf, pairs_path, match_path, feature_path_q, feature_path_ref, overwrite=False, batch=64, device="cuda"):
    """match_features.match_from_paths (:140-186).  conf: an entry of confs_dict["matchers"] (sparse matchers)."""
    feature_path_q, feature_path_ref, match_path = Path(feature_path_q), Path(feature_path_ref), Path(match_path)
    if not feature_path_q.exists():
        raise FileNotFoundError(f"Query feature file {feature_path_q}.")
    if not feature_path_ref.exists():
        raise FileNotFoundError(f"Reference feature file {feature_path_ref}.")
    match_path.parent.mkdir(exist_ok=True, parents=True)
    assert Path(pairs_path).exists(), pairs_path
    pairs = [(q, r) for q, rs in parse_retrieval(pairs_path).items() for r in rs]
    pairs = find_unique_new_pairs(pairs, None if overwrite else match_path)
    if not pairs:
        logger.info("Skipping the matching.")
        return None
    model = dynamic_load(matchers, conf["model"]["name"])(conf["model"]).eval().to(device)
    if not hasattr(model, "match_batch"):
        raise NotImplementedError(f"{conf['model']['name']}: no batched entry (sparse matchers of the hot path have one)")
    loaded = queue.Queue(maxsize=2)
    loader = _PairLoader(pairs, feature_path_q, feature_path_ref, batch, loaded)
    loader.start()
    store = open_store(match_path, "a")
    to_write = queue.Queue(maxsize=4)

    def writer():
        while True:
            item = to_write.get()
            if item is None:
                return
            chunk, m16, s16, counts, ev = item
            ev.synchronize()
            for p, (a, b) in enumerate(chunk):
                n0 = int(counts[2 * p])
                store.write_group(names_to_pair(a, b), {"matches0": m16[p, :n0].numpy().copy(), "matching_scores0": s16[p, :n0].numpy().copy()})
    wt = threading.Thread(target=writer, daemon=True)
    wt.start()
    n_done = 0
    while True:
        item = loaded.get()
        if item is None:
            break
        chunk, host = item
        dev = {k: v.to(device, non_blocking=True) for k, v in host.items()}
        m0, s0 = model.match_batch(dev)
        P, cap = m0.shape
        m16 = torch.empty(P, cap, dtype=torch.int16).pin_memory(); s16 = torch.empty(P, cap, dtype=torch.float16).pin_memory()
        m16.copy_(m0.to(torch.int16), non_blocking=True)       # writer_fn :79-82: .short() / .half()
        s16.copy_(s0.to(torch.float16), non_blocking=True)
        ev = torch.cuda.Event(); ev.record(torch.cuda.current_stream(dev["keypoints"].device))
        to_write.put((chunk, m16, s16, host["counts"], ev))
        n_done += len(chunk)
    to_write.put(None)
    wt.join()
    store.close()
    if loader.error is not None:
        raise loader.error
    logger.info(f"Finished exporting matches ({n_done} pairs).")
    return match_path


def main(conf, pairs, features, export_dir=None, matches=None, features_ref=None, overwrite=False, **kw):
    """match_features.main (:86-114): path or name conventions for the feature / match files"""
    if isinstance(features, Path) or Path(features).exists():
        features_q = Path(features)
        if matches is None:
            raise ValueError("Either provide both features and matches as Path or both as names.")
    else:
        if export_dir is None:
            raise ValueError(f"Provide an export_dir if features is not a file path: {features}.")
        features_q = Path(export_dir, features + ".h5")
        if matches is None:
            matches = Path(export_dir, f'{features}_{conf["output"]}_{Path(pairs).stem}.h5')
    match_from_paths(conf, Path(pairs), Path(matches), features_q, Path(features_ref) if features_ref else features_q, overwrite, **kw)
    return matches
