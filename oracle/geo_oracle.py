"""TEST INFRASTRUCTURE ONLY — numpy restatement of the reference's geometric verification around a given fundamental
matrix (paths relative to /root/reference/SfM), used to check metricsfm_b200/csrc/geo_kernels.cuh.  The RANSAC search
itself lives in OpenCV (cv::findFundamentalMat, un-vendored, version 2.4.13.6 per SfM/CMakeLists.txt:48); the tests
compare against the cv2 build present in this image (statistical parity: its RNG stream cannot be reproduced)."""
from __future__ import annotations

import numpy as np


def f_filter(F: np.ndarray, pt1: np.ndarray, pt2: np.ndarray, th_epi: float = 3.0) -> np.ndarray:
    """GeoVerification::GeoVerificationFundamental(pt1, pt2, F, inliers), src/utils/geo_verification.cc:60-79:
    l1 = F * p1, normalised by sqrt(a^2 + b^2); keep i iff |l1 . p2| < th_epi.  Double arithmetic."""
    F = np.asarray(F, np.float64).reshape(3, 3)
    p1 = np.concatenate([np.asarray(pt1, np.float64), np.ones((len(pt1), 1))], 1)
    p2 = np.concatenate([np.asarray(pt2, np.float64), np.ones((len(pt2), 1))], 1)
    l1 = p1 @ F.T
    n = np.sqrt(l1[:, 0] ** 2 + l1[:, 1] ** 2)
    dis = ((l1 / n[:, None]) * p2).sum(1)
    return np.abs(dis) < th_epi


def ransac_error(F: np.ndarray, pt1: np.ndarray, pt2: np.ndarray) -> np.ndarray:
    """cv::findFundamentalMat's per-point RANSAC error (calib3d fundam.cpp, computeReprojError / computeError): the
    larger of the squared distances of p2 to F p1 and of p1 to F^T p2; a point is an inlier iff err <= th^2."""
    F = np.asarray(F, np.float64).reshape(3, 3)
    p1 = np.concatenate([np.asarray(pt1, np.float64), np.ones((len(pt1), 1))], 1)
    p2 = np.concatenate([np.asarray(pt2, np.float64), np.ones((len(pt2), 1))], 1)
    l2 = p1 @ F.T
    l1 = p2 @ F
    s = (l2 * p2).sum(1)
    d2 = s * s / (l2[:, 0] ** 2 + l2[:, 1] ** 2)
    d1 = s * s / (l1[:, 0] ** 2 + l1[:, 1] ** 2)
    return np.maximum(d1, d2)


def verify_pair(pt1_good, pt2_good, pt1_all, pt2_all, F, th: float = 3.0, min_points: int = 30, min_inliers: int = 30):
    """Decision flow of fine_matching_graph.cc:137-153 around a given F: (ok, stage-A inlier count, stage-B mask)."""
    if len(pt1_good) < min_points:
        return False, 0, np.zeros((len(pt1_all),), bool)
    inl = int((ransac_error(F, pt1_good, pt2_good) <= th * th).sum())
    if inl < min_inliers:
        return False, inl, np.zeros((len(pt1_all),), bool)
    return True, inl, f_filter(F, pt1_all, pt2_all, th)
