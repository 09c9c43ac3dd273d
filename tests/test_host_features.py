"""CPU tests of the host logic around the feature kernel (vvc_intra_b200/features.py): the reference's gate and its
neighbour-CU selection (EL/EncCu.cpp:821-933) on random QT/MTT partitions."""
import numpy as np

import vvc_intra_b200 as vb


def random_partition(rng, W, H, ctu=128):
    """A random legal-looking luma partition: list of dict(x, y, w, h, qt_depth, mt_depth) tiling the picture."""
    out = []

    def rec(x, y, w, h, qt, mt, qt_ok):
        if x >= W or y >= H:
            return
        must = x + w > W or y + h > H or w > 64
        r = rng.random()
        if (must or (r < 0.55 and qt_ok)) and w == h and w > 8 and qt_ok:
            for dy in (0, h // 2):
                for dx in (0, w // 2):
                    rec(x + dx, y + dy, w // 2, h // 2, qt + 1, 0, True)
            return
        if must and w == h and w == 8:                      # 8x8 straddling the boundary: keep the inside part
            pass
        if not must and mt < 3 and w <= 32 and h <= 32 and r > 0.6:
            k = rng.integers(0, 4)
            if k == 0 and h > 4:
                rec(x, y, w, h // 2, qt, mt + 1, False); rec(x, y + h // 2, w, h // 2, qt, mt + 1, False); return
            if k == 1 and w > 4:
                rec(x, y, w // 2, h, qt, mt + 1, False); rec(x + w // 2, y, w // 2, h, qt, mt + 1, False); return
            if k == 2 and h >= 16:
                rec(x, y, w, h // 4, qt, mt + 1, False); rec(x, y + h // 4, w, h // 2, qt, mt + 1, False)
                rec(x, y + 3 * h // 4, w, h // 4, qt, mt + 1, False); return
            if k == 3 and w >= 16:
                rec(x, y, w // 4, h, qt, mt + 1, False); rec(x + w // 4, y, w // 2, h, qt, mt + 1, False)
                rec(x + 3 * w // 4, y, w // 4, h, qt, mt + 1, False); return
        if x + w <= W and y + h <= H:
            out.append(dict(x=x, y=y, w=w, h=h, qt_depth=qt, mt_depth=mt))

    for cy in range(0, H, ctu):
        for cx in range(0, W, ctu):
            rec(cx, cy, ctu, ctu, 0, 0, True)
    return out


def cu_lookup(cus, W, H):
    idx = -np.ones((H, W), np.int32)
    for i, c in enumerate(cus):
        idx[c['y']:c['y'] + c['h'], c['x']:c['x'] + c['w']] = i

    def get_cu(px, py):
        if px < 0 or py < 0 or px >= W or py >= H or idx[py, px] < 0:
            return None
        return cus[idx[py, px]]
    return get_cu


def test_gate():
    assert vb.feature_gate(0, 0, 64, 64, 0)
    assert not vb.feature_gate(0, 0, 4, 4, 2)                 # 4x4 is skipped (EL/EncCu.cpp:842)
    assert not vb.feature_gate(0, 0, 16, 16, 3)               # mtDepth 3 (:837)
    assert not vb.feature_gate(0, 0, 128, 128, 0)             # height < 128 (:835)
    assert not vb.feature_gate(400, 0, 32, 32, 0)             # x + w <= 416, hard-coded (:832-835)
    assert not vb.feature_gate(0, 224, 32, 32, 0)
    assert not vb.feature_gate(0, 0, 32, 32, 0, is_luma=False)


def test_neighbour_selection_on_random_partitions():
    rng = np.random.default_rng(9)
    W, H = 416, 240
    seen = np.zeros(6, int)
    for _ in range(4):
        cus = random_partition(rng, W, H)
        get_cu = cu_lookup(cus, W, H)
        for c in cus:
            x, y, w, h = c['x'], c['y'], c['w'], c['h']
            nbs = vb.select_feature_neighbours(get_cu, x, y, w, h)
            seen[len(nbs)] += 1
            assert len(nbs) <= 5
            # independent restatement of the acceptance rules
            exp = []
            L, U, LU = get_cu(x - 1, y), get_cu(x, y - 1), get_cu(x - 1, y - 1)
            if L:
                exp.append(L)
                LD = get_cu(x - 1, y + L['h'] + 1)
                if LD and LD['y'] <= y + h:
                    exp.append(LD)
            if U:
                exp.append(U)
                RU = get_cu(x + U['w'] + 1, y - 1)
                if RU and RU['x'] < x + w:
                    exp.append(RU)
            if LU and LU['y'] + LU['h'] <= y and LU['x'] + LU['w'] <= x:
                exp.append(LU)
            assert [id(a) for a in nbs] == [id(a) for a in exp]
            if x == 0 and y == 0:
                assert nbs == []
            j = vb.feature_job(x, y, w, h, c['qt_depth'], c['mt_depth'], nbs)
            assert int(j['n_neighbours']) == len(nbs) and int(j['cu']['w']) == w
    assert seen[3:].sum() > 100 and seen[0] >= 4


def test_classifier_decision():
    f = [0] * 27
    f[26] = 0
    assert vb.classifier_decision(f, 0) == 'ETM_INTRA'
    f[26] = 1
    assert vb.classifier_decision(f, 0) is None               # fuzzy / complex blocks are never terminated early (:1197-1203)
    assert vb.classifier_decision(f, 3) == 'ETM_SPLIT_BT_V'
    assert vb.classifier_decision(f, -1) is None and vb.classifier_decision(f, None) is None


def test_training_set_twin_differs_only_on_the_left_down_edge_case():
    """GET_TRAINING_SET (EL/CABACWriter.cpp:515-858) collects the same neighbours as the inference side except for one comparison:
    a left-down CU that starts exactly at the CU's bottom edge is accepted by EncCu (`<=`, :886) and rejected by the writer (`<`, :581)."""
    rng = np.random.default_rng(17)
    cus = random_partition(rng, 416, 240)
    get_cu = cu_lookup(cus, 416, 240)
    differ = same = 0
    for c in cus:
        a = vb.select_feature_neighbours(get_cu, c['x'], c['y'], c['w'], c['h'])
        b = vb.select_feature_neighbours(get_cu, c['x'], c['y'], c['w'], c['h'], training_set=True)
        if a == b:
            same += 1
            continue
        differ += 1
        dropped = [n for n in a if n not in b]
        assert len(dropped) == 1 and b == [n for n in a if n is not dropped[0]]
        assert dropped[0]['y'] == c['y'] + c['h'] and dropped[0]['x'] < c['x']       # the left-down CU on the edge
    assert differ > 5 and same > 5


def random_tree(rng, W, H, ctu=128):
    """A random final luma coding tree: (nodes in coding order with their split, leaf CUs) -- what CABACWriter::coding_tree walks."""
    nodes, leaves = [], []

    def rec(x, y, w, h, qt, mt, qt_ok):
        if x >= W or y >= H:
            return
        must = x + w > W or y + h > H or w > 64
        r = rng.random()
        if (must or (r < 0.55 and qt_ok)) and w == h and w > 8 and qt_ok:
            nodes.append(dict(x=x, y=y, w=w, h=h, split=1))
            for dy in (0, h // 2):
                for dx in (0, w // 2):
                    rec(x + dx, y + dy, w // 2, h // 2, qt + 1, 0, True)
            return
        if not must and mt < 3 and w <= 32 and h <= 32 and r > 0.6:
            k = int(rng.integers(0, 4))
            parts = None
            if k == 0 and h > 4:
                parts, split = [(x, y, w, h // 2), (x, y + h // 2, w, h // 2)], 2
            elif k == 1 and w > 4:
                parts, split = [(x, y, w // 2, h), (x + w // 2, y, w // 2, h)], 3
            elif k == 2 and h >= 16:
                parts, split = [(x, y, w, h // 4), (x, y + h // 4, w, h // 2), (x, y + 3 * h // 4, w, h // 4)], 4
            elif k == 3 and w >= 16:
                parts, split = [(x, y, w // 4, h), (x + w // 4, y, w // 2, h), (x + 3 * w // 4, y, w // 4, h)], 5
            if parts:
                nodes.append(dict(x=x, y=y, w=w, h=h, split=split))
                for (px, py, pw, ph) in parts:
                    rec(px, py, pw, ph, qt, mt + 1, False)
                return
        if x + w <= W and y + h <= H:
            nodes.append(dict(x=x, y=y, w=w, h=h, split=0))
            leaves.append(dict(x=x, y=y, w=w, h=h, qt_depth=qt, mt_depth=mt))

    for cy in range(0, H, ctu):
        for cx in range(0, W, ctu):
            rec(cx, cy, ctu, ctu, 0, 0, True)
    return nodes, leaves


def test_training_set_dump_host_logic(tmp_path):
    """GET_TRAINING_SET (EL/CABACWriter.cpp:515-858): gate, depths taken from the top-left leaf, `<` in the left-down test, labels and the four files;
    the features come from the oracle here and from the feature kernel in tests/test_gpu_parity.py."""
    from oracle import oracle_py as O
    from make_golden import synth_yuv
    from vvc_intra_b200 import training_set as T
    rng = np.random.default_rng(11)
    nodes, leaves = random_tree(rng, 416, 240)
    get_cu = cu_lookup(leaves, 416, 240)
    jobs, labels = T.training_jobs(nodes, get_cu)
    assert len(jobs) > 100 and set(labels.tolist()) >= {0, 1, 2, 3}
    # an internal node carries the depths of the leaf at its top-left sample, a leaf at multi-type depth 3 is not written, 4x4 never
    for j, lab in zip(jobs, labels):
        c = j['cu']
        leaf = get_cu(int(c['x']), int(c['y']))
        assert (int(c['qt_depth']), int(c['mt_depth'])) == (leaf['qt_depth'], leaf['mt_depth'])
        assert not (lab == 0 and leaf['mt_depth'] == 3) and not (c['w'] == 4 and c['h'] == 4) and j['n_neighbours'] >= 3
    Y = synth_yuv(416, 240, 8)[0].astype(np.int16)
    res = O.features_batch(Y, jobs)
    n = T.write_training_set(str(tmp_path), res['f'], labels)
    n += T.write_training_set(str(tmp_path), res['f'][:5], labels[:5])           # the files are appended to (CABACWriter.h:60-63: "a+b")
    data = np.fromfile(tmp_path / 'Data_Partition.dat', '<i4').reshape(-1, 26)
    assert n == len(jobs) + 5 == len(data) and np.array_equal(data[:len(jobs)], res['f'][:, :26])
    assert np.array_equal(np.fromfile(tmp_path / 'Data_Termination.dat', '<i4').reshape(-1, 26), data)
    part = np.fromfile(tmp_path / 'Label_Partition.dat', '<i4')
    term = np.fromfile(tmp_path / 'Label_Termination.dat', '<i4')
    assert np.array_equal(part[:len(jobs)], labels) and np.array_equal(term, (part != 0).astype(np.int32))
    assert data[0, 0] == jobs[0]['cu']['h'] and data[0, 1] == jobs[0]['cu']['w']             # Training_set[0] = height, [1] = width (:797-798)
