"""On-disk formats around the matching path: libmsfm_store.so against the pure-Python
restatement of the reference's writers/readers (oracle/store_oracle.py), byte for byte.  CPU only."""
import os

import numpy as np
import pytest

from metricsfm_b200 import store, synth
from oracle import store_oracle as so


@pytest.fixture(scope="module", autouse=True)
def _built():
    from metricsfm_b200 import build
    build.build_native()
    build.build_host_libs()


def _feature_case(rng, n, dtype):
    xy = rng.uniform(0, 4000, size=(n, 2)).astype(np.float32)
    desc = (rng.uniform(0, 1, size=(n, 128)).astype(np.float32) if dtype == np.float32
            else rng.integers(0, 256, size=(n, 128), dtype=np.uint8))
    return dict(rows=3000, cols=4001, zoom_ratio=0.5, f_mm=24.0, f_pixel=3100.5, gps_latitude=32.1, gps_longitude=118.9,
                maker="DJI", model="FC6310 mk2", xy_pixel=xy, desc=desc)


@pytest.mark.parametrize("n,dtype", [(0, np.float32), (1, np.float32), (257, np.float32), (300, np.uint8)])
def test_feature_file_bytes_and_roundtrip(tmp_path, n, dtype):
    rng = np.random.default_rng(n + 7)
    case = _feature_case(rng, n, dtype)
    path = store.feature_path(str(tmp_path), 12)
    assert path == str(tmp_path) + "//12_feature"
    store.feature_write(path, **case)
    blob = open(path, "rb").read()
    assert blob == so.feature_bytes(**case)                      # writer == Database::WriteoutImageFeature
    got = store.feature_read(path)                                # reader == Database::ReadinImageFeatures
    exp = so.feature_parse(blob)
    assert (got["info"].rows, got["info"].cols, got["info"].num_pts) == (exp["rows"], exp["cols"], n)
    assert got["maker"] == "DJI" and got["model"] == "FC6310 mk2"
    np.testing.assert_array_equal(got["xy"], exp["xy"])
    np.testing.assert_array_equal(got["desc"], case["desc"])
    # keypoints are stored centred with the reference's double arithmetic (odd image width: cols / 2.0 = 2000.5)
    np.testing.assert_array_equal(got["xy"][:, 0], (case["xy_pixel"][:, 0].astype(np.float64) - 2000.5).astype(np.float32))


def test_feature_file_errors(tmp_path):
    with pytest.raises(store.StoreError):
        store.feature_stat(str(tmp_path / "nope"))
    p = tmp_path / "3_feature"
    case = _feature_case(np.random.default_rng(0), 50, np.float32)
    store.feature_write(str(p), **case)
    blob = p.read_bytes()
    p.write_bytes(blob[:-100])                                    # truncated descriptor block
    with pytest.raises(store.StoreError):
        store.feature_stat(str(p))


def test_match_file_append_read_and_recover(tmp_path):
    fold = str(tmp_path)
    rng = np.random.default_rng(1)
    records = [(5, rng.integers(0, 9000, size=(40, 2)).astype(np.int32)), (2, np.empty((0, 2), np.int32)),
               (9, rng.integers(0, 9000, size=(1, 2)).astype(np.int32)), (7, rng.integers(0, 9000, size=(333, 2)).astype(np.int32))]
    expect = b""
    for idx2, pairs in records:
        store.match_append(fold, 3, idx2, pairs)
        expect += so.match_record_bytes(idx2, pairs)
    assert open(fold + "//3_match", "rb").read() == expect       # appended records == WriteOutMatches
    ids, lists = store.match_read(fold, 3)                        # == Graph::QueryMatch
    oids, olists = so.match_parse(expect)
    assert ids.tolist() == oids == [5, 9, 7]                      # the empty list left no record
    for a, b in zip(lists, olists):
        np.testing.assert_array_equal(a, b)
    g = store.graph_recover(fold, 10, [3, 4])                     # image 4 has no file: skipped like the reference
    exp = np.zeros((10, 10), np.int32)
    exp[3, 5], exp[3, 9], exp[3, 7] = 40, 1, 333
    np.testing.assert_array_equal(g, exp)


def test_match_index_resume(tmp_path):
    fold = str(tmp_path)
    np.testing.assert_array_equal(store.match_index_missing(fold, 5), np.arange(5))   # no file: everything is missing
    for i in (3, 0, 3):
        store.match_index_append(fold, i)
    assert open(fold + "//match_index.txt").read() == "3\n0\n3\n"
    np.testing.assert_array_equal(store.match_index_missing(fold, 5), [1, 2, 4])
    assert so.missing_from_index_text("3\n0\n3\n", 5) == [1, 2, 4]


def test_graph_matching_text(tmp_path):
    fold = str(tmp_path)
    g = np.random.default_rng(2).integers(0, 5000, size=(7, 7)).astype(np.int32)
    store.graph_write(fold, g)
    assert open(fold + "//graph_matching.txt", "rb").read() == so.graph_text(g)
    np.testing.assert_array_equal(store.graph_read(fold, 7), g)


def test_pair_lists_all_and_priori(tmp_path):
    offs, lst = store.pairs_all(6)
    assert store._adjacency(offs, lst) == so.pairs_all(6)
    # aerial block: jittered flight grid + two exposures at (almost) the same position (redundancy filter)
    rng = np.random.default_rng(3)
    gx, gy = np.meshgrid(np.arange(15), np.arange(14))
    xy = np.stack([gx.ravel() * 37.0, gy.ravel() * 53.0], 1) + rng.normal(0, 3.0, size=(210, 2))
    xy[100] = xy[99] + 0.2
    for knn in (50, 8, 0):
        offs, lst = store.pairs_priori_xy(xy, knn)
        adj = store._adjacency(offs, lst)
        assert adj == so.pairs_priori_xy(xy, knn)
    assert adj == [[] for _ in range(210)]
    offs, lst = store.pairs_priori_xy(xy, 8)
    adj = store._adjacency(offs, lst)
    assert sum(1 for a in adj if not a) >= 1 and max(len(a) for a in adj) == 8
    # init_match_graph.txt round trip
    fold = str(tmp_path)
    store.init_graph_write(fold, offs, lst, 209)
    assert open(fold + "//init_match_graph.txt", "rb").read() == so.init_graph_text(adj, 209)
    o2, l2, id_last = store.init_graph_read(fold)
    assert id_last == 209
    np.testing.assert_array_equal(o2, offs)
    np.testing.assert_array_equal(l2, lst)


def test_store_and_graph_libraries_export_their_headers():
    """Every function the two headers declare is exported (no compute calls here)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for header, symbols, lib in (("msfm_store.h", store.STORE_SYMBOLS, store.lib()), ("msfm_graph.h", store.GRAPH_SYMBOLS, store.graph_lib())):
        text = open(os.path.join(root, "include", header)).read()
        declared = sorted(set(re.findall(r"^int (msfm_\w+)\(", text, flags=re.M)))
        assert declared == sorted(symbols)
        for name in declared:
            assert hasattr(lib, name)


def test_product_tree_does_not_touch_the_store_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "metricsfm_b200")):
        for f in files:
            if f.endswith((".py", ".cc", ".cu", ".cuh", ".h")):
                assert "store_oracle" not in open(os.path.join(dirpath, f), errors="ignore").read(), f


def test_bow_retrieval_route_matches_restatement():
    """InitialMatchingGraph::match_graph_feature's candidate generation (inverted-file similarity -> top-k hypotheses ->
    word-collision matches) against the line-by-line restatement in oracle/store_oracle.py, including the quirks of the
    reference's keep_unique helpers."""
    from oracle import store_oracle as so
    rng = np.random.default_rng(5)
    num_words, n = 3000, 40
    shared = rng.integers(0, num_words, size=600)
    words = []
    for i in range(n):
        own = rng.integers(0, num_words, size=rng.integers(300, 900))
        take = shared[rng.random(len(shared)) < (0.5 if i % 3 == 0 else 0.1)]
        w = np.concatenate([own, take]).astype(np.int32)
        rng.shuffle(w)
        words.append(w)
    words[7] = np.zeros((0,), np.int32)                      # an image without words
    words[8] = np.array([5, 5, 5, 9], np.int32)               # duplicates only + the largest id: nothing survives
    sim = store.similarity_invfile(words, num_words)
    np.testing.assert_array_equal(sim, so.similarity_invfile(words, num_words))
    assert sim.sum() > 0 and np.array_equal(sim, sim.T) and sim[7].sum() == 0 and sim[8].sum() == 0
    for k in (0, 5):
        offs, lst = store.pairs_similarity_topk(sim, k)
        exp = so.pairs_similarity_topk(sim, k)
        for i in range(n):
            assert list(lst[offs[i]:offs[i + 1]]) == exp[i]
    # hand-checked quirk: ids sorted = [1, 2, 2, 3, 4, 9] -> keep_unique_vector keeps 3 and 4 (1 is the first run, 9 the last)
    assert so.keep_unique_vector([2, 9, 1, 4, 2, 3]) == [3, 4]
    tot = 0
    for a, b in [(0, 3), (3, 6), (1, 2), (7, 0), (8, 0), (9, 12)]:
        got = store.word_matches(words[a], words[b])
        exp = np.array(so.word_matches(words[a], words[b]), np.int32).reshape(-1, 2)
        np.testing.assert_array_equal(got, exp)
        for p1, p2 in got:
            assert words[a][p1] == words[b][p2]
        tot += len(got)
    assert tot > 30
