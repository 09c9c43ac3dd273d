"""GET_TRAINING_SET (EL/CABACWriter.cpp:515-858): the fork dumps, from the final bitstream writer, the 26 classifier inputs of every node of the
chosen luma coding tree together with the split the encoder chose there -- the training data of the partition / termination classifiers.

The pixel work is the same feature kernel the FAST_ALGORITHM path uses (vvcb_features_eval); what differs is host logic, restated here:
  * the walk is CABACWriter::coding_tree's: every node of the final tree in coding order, internal nodes included;
  * a node's depths are those of the CU covering its top-left sample (cs.getCU( currArea ), :477 -- for an internal node that is a descendant leaf);
  * gate :530-541: height < 128, inside the hard-coded 416x240, not (no split at multi-type depth 3), not 4x4;
  * neighbour selection as in EncCu but with `<` for the left-down CU (:581), at least three neighbours (:634);
  * labels :823-848: partition = 0 for no split else the PartSplit value (1 QT, 2 BT_H, 3 BT_V, 4 TT_H, 5 TT_V), termination = 0 / 1;
  * files: Data_Partition.dat / Data_Termination.dat (26 int32 per record), Label_Partition.dat / Label_Termination.dat (1 int32), appended (CABACWriter.h:60-63)."""
import os

import numpy as np

from .engine import FEAT_JOB_DTYPE
from .features import VIDEO_WIDTH, VIDEO_HEIGHT, select_feature_neighbours, feature_job

SPLIT_NONE, SPLIT_QT, SPLIT_BT_H, SPLIT_BT_V, SPLIT_TT_H, SPLIT_TT_V = 0, 1, 2, 3, 4, 5


def training_gate(w, h, x, y, split, mt_depth):
    if not (h < 128 and x + w <= VIDEO_WIDTH and y + h <= VIDEO_HEIGHT):
        return False
    if split == SPLIT_NONE and mt_depth == 3:
        return False
    return not (h == 4 and w == 4)


def training_jobs(nodes, get_cu):
    """nodes: dicts (x, y, w, h, split) of the final luma tree in coding order; get_cu(px, py): leaf CU dict covering a luma position.
    Returns (jobs, labels): FEAT_JOB_DTYPE array and the split of each job's node, for the nodes the reference writes a record for."""
    jobs, labels = [], []
    for nd in nodes:
        cu = get_cu(nd['x'], nd['y'])
        if cu is None or not training_gate(nd['w'], nd['h'], nd['x'], nd['y'], nd['split'], cu['mt_depth']):
            continue
        nbs = select_feature_neighbours(get_cu, nd['x'], nd['y'], nd['w'], nd['h'], training_set=True)
        if len(nbs) < 3:
            continue
        jobs.append(feature_job(nd['x'], nd['y'], nd['w'], nd['h'], cu['qt_depth'], cu['mt_depth'], nbs))
        labels.append(nd['split'])
    return np.array(jobs, FEAT_JOB_DTYPE), np.array(labels, np.int32)


def write_training_set(out_dir, features, labels):
    """Appends the records to the reference's four files.  features: (n, >= 26) int32 (vvcb_feat_result.f), labels: the nodes' splits."""
    feats = np.ascontiguousarray(np.asarray(features, '<i4')[:, :26])
    part = np.asarray(labels, '<i4')
    term = (part != SPLIT_NONE).astype('<i4')
    os.makedirs(out_dir, exist_ok=True)
    for name, data in (('Data_Partition.dat', feats), ('Label_Partition.dat', part), ('Data_Termination.dat', feats), ('Label_Termination.dat', term)):
        with open(os.path.join(out_dir, name), 'ab') as f:
            f.write(data.tobytes())
    return len(part)


def dump_training_set(engine, nodes, get_cu, out_dir):
    """The whole dump for one picture whose (LMCS-mapped) original luma was given to engine.frame_begin: features on the device, files on the host."""
    jobs, labels = training_jobs(nodes, get_cu)
    if not len(jobs):
        return 0
    res = engine.features_eval(jobs)
    return write_training_set(out_dir, res['f'], labels)
