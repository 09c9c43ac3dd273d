"""Host logic around the FAST_ALGORITHM feature kernel (vvcb_features_eval): what the reference decides on the host
before and after the pixel work.  Restates EL/EncCu.cpp:821-933 (gate and neighbour selection) and :1172-1216 (how the
classifier's answer is applied).  The classifier itself is a CPython call into a pickled model the repository does not
ship (BIN/TEST.py:14-24); `classifier_decision` takes the predicted class from the caller."""
import numpy as np

from .engine import FEAT_JOB_DTYPE

# hard-coded in the reference (EL/EncCu.cpp:832-833), whatever the real picture size
VIDEO_WIDTH, VIDEO_HEIGHT = 416, 240
# GetPartition result -> split tried exclusively (EL/EncCu.cpp:1172-1195)
PARTITION_OF_CLASS = ('ETM_INTRA', 'ETM_SPLIT_QT', 'ETM_SPLIT_BT_H', 'ETM_SPLIT_BT_V', 'ETM_SPLIT_TT_H', 'ETM_SPLIT_TT_V')


def feature_gate(x, y, w, h, mt_depth, is_luma=True):
    """True when the reference computes features for this CU (EL/EncCu.cpp:821-845)."""
    if not is_luma:
        return False
    if not (h < 128 and x + w <= VIDEO_WIDTH and y + h <= VIDEO_HEIGHT):
        return False
    if mt_depth == 3:
        return False
    return not (h == 4 and w == 4)


def select_feature_neighbours(get_cu, x, y, w, h, training_set=False):
    """Neighbour CUs in the order the reference collects them (EL/EncCu.cpp:852-933).

    get_cu(px, py) -> None or a dict(x, y, w, h, qt_depth, mt_depth): the coding structure's CU covering luma
    position (px, py) (tempCS->getCU).  Returns the list of accepted neighbours (valid_num = its length).
    training_set=True follows the GET_TRAINING_SET twin that dumps the same features from the bitstream writer
    (EL/CABACWriter.cpp:515-858): identical except that it accepts the left-down CU only when it starts strictly above the
    CU's bottom edge (`<` at EL/CABACWriter.cpp:581 against `<=` at EL/EncCu.cpp:886)."""
    out = []
    left = get_cu(x - 1, y)
    up = get_cu(x, y - 1)
    left_up = get_cu(x - 1, y - 1)
    if left is not None:
        out.append(left)
        left_down = get_cu(x - 1, y + left['h'] + 1)             # offset(-1, cuLeft->lheight() + 1), :873
        if left_down is not None and (left_down['y'] < y + h if training_set else left_down['y'] <= y + h):    # :886 / CABACWriter.cpp:581
            out.append(left_down)
    if up is not None:
        out.append(up)
        right_up = get_cu(x + up['w'] + 1, y - 1)                # offset(cuUp->lwidth() + 1, -1), :894
        if right_up is not None and right_up['x'] < x + w:       # :897
            out.append(right_up)
    if left_up is not None:
        if not (left_up['y'] + left_up['h'] > y or left_up['x'] + left_up['w'] > x):   # :910
            out.append(left_up)
    return out


def feature_job(x, y, w, h, qt_depth, mt_depth, neighbours):
    j = np.zeros(1, FEAT_JOB_DTYPE)[0]
    j['cu'] = (x, y, w, h, qt_depth, mt_depth)
    j['n_neighbours'] = len(neighbours)
    for i, c in enumerate(neighbours):
        j['nb'][i] = (c['x'], c['y'], c['w'], c['h'], c['qt_depth'], c['mt_depth'])
    return j


def classifier_decision(features, predicted_class):
    """What xCompressCU does with GetPartition's answer (EL/EncCu.cpp:1172-1216): returns the only test mode to keep,
    or None when the normal search runs (classifier not applicable: f26 >= 1 and 'no split' predicted, :1197)."""
    if predicted_class is None or not 0 <= predicted_class < len(PARTITION_OF_CLASS):
        return None
    if features[26] >= 1 and predicted_class == 0:
        return None
    return PARTITION_OF_CLASS[predicted_class]


def build_sweep_feature_jobs(width, height, ctu=128):
    """Feature jobs of the exhaustive sweep of one picture (bench / full-size tests): every candidate CU of every 64x64 root
    except the 4x4 ones (EL/EncCu.cpp:842), with the neighbour CUs a uniform partition into blocks of the CU's own size would
    offer -- left, left-down, up and left-up where they lie inside the picture (a right-up block of the same size starts at
    x + w and fails the reference's `< x + w` test, :897).  The picture-size gate of `feature_gate` (the reference's hard-coded
    416x240) is the caller's business and not applied here."""
    from .partition import frame_candidates
    c = frame_candidates(width, height)
    c = c[~((c[:, 2] == 4) & (c[:, 3] == 4))]
    x, y, w, h = (c[:, i].astype(np.int64) for i in range(4))
    side = np.maximum(w, h)
    lg = np.log2(side).astype(np.int64)
    qt = (np.log2(ctu).astype(np.int64) - lg).astype(np.uint8)
    mt = np.minimum(3, (lg - np.log2(w).astype(np.int64)) + (lg - np.log2(h).astype(np.int64))).astype(np.uint8)
    jobs = np.zeros(len(c), FEAT_JOB_DTYPE)
    jobs['cu']['x'], jobs['cu']['y'], jobs['cu']['w'], jobs['cu']['h'] = x, y, w, h
    jobs['cu']['qt_depth'], jobs['cu']['mt_depth'] = qt, mt
    has_left, has_up = x >= w, y >= h
    cand = [(has_left, x - w, y), (has_left & (y + 2 * h <= height), x - w, y + h), (has_up, x, y - h), (has_left & has_up, x - w, y - h)]
    n = np.zeros(len(c), np.int64)
    idx = np.arange(len(c))
    for ok, nx, ny in cand:
        sel = idx[ok]
        k = n[sel]
        nb = jobs['nb']
        nb['x'][sel, k], nb['y'][sel, k], nb['w'][sel, k], nb['h'][sel, k] = nx[sel], ny[sel], w[sel], h[sel]
        nb['qt_depth'][sel, k], nb['mt_depth'][sel, k] = qt[sel], mt[sel]
        n[sel] += 1
    jobs['n_neighbours'] = n
    return jobs
