"""TEST INFRASTRUCTURE ONLY — pure-Python restatement of the reference's on-disk formats around the matching path and of
FineMatchingGraph::BuildMatchGraph's bookkeeping, used to check metricsfm_b200/host/msfm_store.cc / msfm_graph.cc byte
for byte.  Paths relative to /root/reference/SfM.  Nothing under metricsfm_b200/ may import this module."""
from __future__ import annotations

import os
import struct

import numpy as np


def join(fold: str, name: str) -> str:
    return fold + "//" + name  # database.cc:360


# ---- Database::WriteoutImageFeature, src/database.cc:490-541 -------------------------------------------------------
def feature_bytes(*, rows, cols, zoom_ratio, f_mm, f_pixel, gps_latitude, gps_longitude, maker: str, model: str,
                  xy_pixel: np.ndarray, desc: np.ndarray) -> bytes:
    out = struct.pack("<ii5f", rows, cols, zoom_ratio, f_mm, f_pixel, gps_latitude, gps_longitude)
    mk, md = maker.encode("latin-1"), model.encode("latin-1")
    out += struct.pack("<i", len(mk)) + mk + struct.pack("<i", len(md)) + md
    xy = np.asarray(xy_pixel, np.float32).reshape(-1, 2)
    out += struct.pack("<i", xy.shape[0])
    centred = np.empty_like(xy)
    centred[:, 0] = (xy[:, 0].astype(np.float64) - cols / 2.0).astype(np.float32)  # :524
    centred[:, 1] = (xy[:, 1].astype(np.float64) - rows / 2.0).astype(np.float32)  # :525
    out += centred.tobytes()
    type_code = {np.dtype(np.float32): 5, np.dtype(np.uint8): 0}[desc.dtype]  # CV_32FC1 / CV_8UC1
    out += struct.pack("<iii", desc.shape[0], desc.shape[1], type_code) + np.ascontiguousarray(desc).tobytes()
    return out


# ---- Database::ReadinImageFeatures, src/database.cc:352-423 ----------------------------------------------------------
def feature_parse(blob: bytes) -> dict:
    o = 0
    rows, cols, zoom, f_mm, f_px, lat, lon = struct.unpack_from("<ii5f", blob, o); o += 28
    (n,) = struct.unpack_from("<i", blob, o); o += 4
    maker = blob[o:o + n].decode("latin-1"); o += n
    (n,) = struct.unpack_from("<i", blob, o); o += 4
    model = blob[o:o + n].decode("latin-1"); o += n
    (npts,) = struct.unpack_from("<i", blob, o); o += 4
    xy = np.frombuffer(blob, np.float32, 2 * npts, o).reshape(-1, 2).copy(); o += 8 * npts
    drows, dcols, dtype = struct.unpack_from("<iii", blob, o); o += 12
    np_dtype = {5: np.float32, 0: np.uint8}[dtype]
    desc = np.frombuffer(blob, np_dtype, drows * dcols, o).reshape(drows, dcols).copy()
    return dict(rows=rows, cols=cols, zoom_ratio=zoom, f_mm=f_mm, f_pixel=f_px, gps_latitude=lat, gps_longitude=lon, maker=maker,
                model=model, xy=xy, desc=desc)


# ---- FineMatchingGraph::WriteOutMatches, src/graph/fine_matching_graph.cc:247-272 --------------------------------------
def match_record_bytes(idx2: int, pairs: np.ndarray) -> bytes:
    pairs = np.asarray(pairs, np.int32).reshape(-1, 2)
    if pairs.shape[0] == 0:
        return b""  # :250-253 returns before opening the file
    return struct.pack("<ii", idx2, pairs.shape[0]) + pairs.tobytes()


# ---- Graph::QueryMatch, src/graph.cc:92-121 -----------------------------------------------------------------------------
def match_parse(blob: bytes):
    ids, lists, o = [], [], 0
    while o + 4 <= len(blob):
        idx2, n = struct.unpack_from("<ii", blob, o); o += 8
        lists.append(np.frombuffer(blob, np.int32, 2 * n, o).reshape(-1, 2).copy()); o += 8 * n
        ids.append(idx2)
    return ids, lists


# ---- CheckMissingMatchingFile, src/graph/fine_matching_graph.cc:209-244 ---------------------------------------------------
def missing_from_index_text(text: str | None, num_imgs: int):
    if text is None:
        return list(range(num_imgs))
    done = {int(t) for t in text.split()}
    return [i for i in range(num_imgs) if i not in done]


# ---- WriteOutMatchGraph, src/graph/fine_matching_graph.cc:275-292 ---------------------------------------------------------
def graph_text(graph: np.ndarray) -> bytes:
    return b"".join(b"".join(b"%d " % int(v) for v in row) + b"\n" for row in np.asarray(graph))


# ---- WriteOutInitMatchGraph, src/graph/initial_matching_graph.cc:324-344 ---------------------------------------------------
def init_graph_text(adj, id_last: int) -> bytes:
    out = b"%d\n%d\n" % (len(adj), id_last)
    for partners in adj:
        out += b"%d " % len(partners) + b"".join(b"%d " % int(j) for j in partners) + b"\n"
    return out


# ---- matching_type "all" / "priori xy", src/graph/initial_matching_graph.cc:55-64, 114-162 ---------------------------------
def pairs_all(num_imgs: int):
    return [[j for j in range(num_imgs) if j != i] for i in range(num_imgs)]


def pairs_priori_xy(xy: np.ndarray, knn: int):
    xy = np.asarray(xy, np.float64).reshape(-1, 2)
    n = xy.shape[0]
    order = sorted(range(n), key=lambda i: (xy[i, 0] + xy[i, 1], i))
    redundant = [False] * n
    if n:
        dis_pre = xy[order[0], 0] + xy[order[0], 1] - 100.0
        for i in order:
            d = xy[i, 0] + xy[i, 1]
            if abs(d - dis_pre) < 1.0:
                redundant[i] = True
            else:
                dis_pre = d
    k = min(knn, n // 10)
    adj = []
    for i in range(n):
        if redundant[i]:
            adj.append([])
            continue
        info = [(abs(xy[i, 0] - xy[j, 0]) + abs(xy[i, 1] - xy[j, 1]), j) for j in range(n) if j != i and not redundant[j]]
        info.sort()
        adj.append([j for _, j in info[:k]])
    return adj


# ---- FineMatchingGraph::BuildMatchGraph bookkeeping, src/graph/fine_matching_graph.cc:40-194 ----------------------------------
def build_match_graph_files(fold: str, adj, match_fn, *, min_good: int = 0) -> None:
    """Reference control flow with the matcher abstracted: match_fn(idx1, idx2) -> (ok, pairs_all [n,2], good [n]).
    Writes <idx1>_match, match_index.txt and graph_matching.txt exactly as the reference would for those lists."""
    n = len(adj)
    index_path = join(fold, "match_index.txt")
    text = open(index_path).read() if os.path.exists(index_path) else None
    missing = missing_from_index_text(text, n)
    if not missing:
        return
    graph = np.zeros((n, n), np.int32)
    for idx in (i for i in range(n) if i not in set(missing)):  # RecoverMatchingGraph :294-330
        p = join(fold, f"{idx}_match")
        if os.path.exists(p):
            ids, lists = match_parse(open(p, "rb").read())
            for i2, l in zip(ids, lists):
                graph[idx, i2] = len(l)
    for idx1 in missing:
        for idx2 in adj[idx1]:
            ok, pairs, good = match_fn(idx1, idx2)
            if not ok or int(np.sum(good)) < min_good:
                continue
            rec = match_record_bytes(idx2, pairs)
            if rec:
                with open(join(fold, f"{idx1}_match"), "ab") as f:
                    f.write(rec)
            graph[idx1, idx2] = len(pairs)
        with open(index_path, "a") as f:
            f.write(f"{idx1}\n")
    with open(join(fold, "graph_matching.txt"), "wb") as f:
        f.write(graph_text(graph))


# ---------------------------------------------------------------------------------------------------- BoW retrieval route
def keep_unique_vector(data):
    """math::keep_unique_vector, SfM/src/utils/basic_funcs.h:126-151, line by line (including what its loop really does:
    the first run is never flagged unique, the last run is never flushed)."""
    data = sorted(int(x) for x in data)
    out = []
    if not data:
        return out
    v, is_unique = data[0], True
    for x in data:
        if x != v:
            if is_unique:
                out.append(v)
            v, is_unique = x, True
        else:
            is_unique = False
    return out


def keep_unique_idx_vector(data):
    """math::keep_unique_idx_vector, SfM/src/utils/basic_funcs.cc:380-406 (stable sort where the reference's std::sort
    leaves the order of equal words unspecified)."""
    d = sorted(((i, int(w)) for i, w in enumerate(data)), key=lambda t: t[1])
    out = []
    if not d:
        return out
    is_unique, v = True, d[0][1]
    for i, w in d:
        if w != v:
            if is_unique:
                out.append(i)
            v, is_unique = w, True
        else:
            is_unique = False
    return out


def similarity_invfile(words_per_image, num_words):
    """SimilarityGraph::SimilarityGraphInvFile + GenerateInvertedFile, SfM/src/graph/similarity_graph.cc:47-117."""
    import numpy as np
    n = len(words_per_image)
    inverted = [[] for _ in range(num_words)]
    for i, w in enumerate(words_per_image):
        if len(w):
            for wid in keep_unique_vector(w):
                inverted[wid].append(i)
    th_bin_size = num_words // 100
    sim = np.zeros((n, n), np.float32)
    for b in inverted:
        if b and len(b) > th_bin_size:
            continue
        for m in range(len(b) - 1):
            for k in range(m + 1, len(b)):
                sim[b[m], b[k]] += 1
                sim[b[k], b[m]] += 1
    return sim


def pairs_similarity_topk(sim, th_num_match=0):
    """initial_matching_graph.cc:166-168 and :212-231 (ties: lower index first)."""
    n = sim.shape[0]
    if th_num_match <= 0:
        th_num_match = min(min(max(200, n // 10), n - 1), 500)
    out = []
    for i in range(n):
        cand = [(j, float(sim[i, j])) for j in range(n) if j != i and not sim[i, j] < 0]
        cand.sort(key=lambda t: -t[1])
        out.append([j for j, _ in cand[:th_num_match]])
    return out


def word_matches(words1, words2):
    """initial_matching_graph.cc:190-201 (pt_word_map) and :239-251 (collisions, ascending word id like std::map)."""
    def pt_word_map(w):
        m = {}
        for idx in keep_unique_idx_vector(w):
            m.setdefault(int(w[idx]), idx)      # std::map::insert keeps the first entry of a key
        return m
    m1, m2 = pt_word_map(words1), pt_word_map(words2)
    return [(m1[w], m2[w]) for w in sorted(m1) if w in m2]

