"""CPU: known answers of the NMS oracle (oracle/nms_oracle.py) - analytic rotated-rectangle overlaps and the greedy
suppression semantics of iou3d_nms.cpp:116-131 / model_nms_utils.py:6-27."""
import numpy as np

from oracle import nms_oracle as no


def box(x, y, dx, dy, a, z=0.0, dz=1.5):
    return np.array([x, y, z, dx, dy, dz, a], dtype=np.float64)


def test_overlap_known_answers():
    a = box(0, 0, 4, 2, 0)
    assert abs(no.bev_overlap(a, a) - 8.0) < 1e-12
    assert no.bev_overlap(a, box(10, 0, 4, 2, 0)) == 0.0
    assert abs(no.bev_overlap(a, box(2, 0, 4, 2, 0)) - 4.0) < 1e-12               # half overlap along x
    assert abs(no.bev_overlap(a, box(0, 0, 4, 2, np.pi)) - 8.0) < 1e-9            # heading + pi is the same rectangle
    # unit square vs the same square turned by 45 degrees: a regular octagon of area 2 (sqrt(2) - 1) * ... = 4 (sqrt(2) - 1) / 2 * 2
    sq = box(0, 0, 2, 2, 0)
    octagon = 8.0 * (np.sqrt(2.0) - 1.0)
    assert abs(no.bev_overlap(sq, box(0, 0, 2, 2, np.pi / 4)) - octagon) < 1e-9
    # containment
    assert abs(no.bev_overlap(box(0, 0, 10, 10, 0.3), box(1, -1, 2, 1, 1.1)) - 2.0) < 1e-9
    iou = no.boxes_iou_bev(np.stack([a, box(2, 0, 4, 2, 0)]), np.stack([a]))
    assert abs(iou[0, 0] - 1.0) < 1e-12 and abs(iou[1, 0] - 4.0 / 12.0) < 1e-12


def test_greedy_suppression_and_class_agnostic_wrapper():
    boxes = np.stack([box(0, 0, 4, 2, 0), box(0.5, 0, 4, 2, 0), box(1.0, 0, 4, 2, 0), box(20, 0, 4, 2, 0.5), box(20.2, 0, 4, 2, 0.5)])
    scores = np.array([0.9, 0.8, 0.95, 0.3, 0.6])
    keep = no.nms(boxes, scores, 0.5)
    # 2 (0.95) suppresses 1 (IoU 0.75) and 0 (IoU 0.6); 4 suppresses 3
    assert keep.tolist() == [2, 4]
    # a chain: with a higher threshold 0 survives because only 1 overlaps it enough and 1 is already removed
    keep = no.nms(boxes, scores, 0.7)
    assert keep.tolist() == [2, 0, 4]
    sel = no.class_agnostic_nms(scores, boxes, 0.5, pre_max=4, post_max=1, score_thresh=0.5)
    assert sel.tolist() == [2]
    sel = no.class_agnostic_nms(scores, boxes, 0.5, pre_max=10, post_max=10, score_thresh=0.85)
    assert sel.tolist() == [2]                       # 0 passes the score mask but is suppressed by 2
    assert no.class_agnostic_nms(scores, boxes, 0.5, 10, 10, score_thresh=0.99).shape == (0,)
    # ties: lower index first
    assert no.nms(boxes[:2], np.array([0.5, 0.5]), 0.5).tolist() == [0]


def test_reference_procedure_restatement_agrees_with_the_exact_overlap():
    """ref_overlap_f32 follows the reference kernel's procedure: away from its 1e-2 m corner margin it is the exact area
    to float32 rounding; a corner within the margin of the other box's side shows the kernel's known error."""
    a = box(0, 0, 4, 2, 0)
    assert abs(no.ref_overlap_f32(a, box(2, 0, 4, 2, 0.0)) - 4.0) < 1e-5
    sq = box(0, 0, 2, 2, 0)
    assert abs(no.ref_overlap_f32(sq, box(0, 0, 2, 2, np.pi / 4)) - 8.0 * (np.sqrt(2.0) - 1.0)) < 1e-5
    assert no.ref_overlap_f32(a, box(10, 0, 4, 2, 0.3)) == 0.0
    g = np.random.default_rng(3)
    worst = 0.0
    for _ in range(60):
        b1 = box(g.uniform(-2, 2), g.uniform(-2, 2), g.uniform(2, 5), g.uniform(1, 3), g.uniform(-3, 3))
        b2 = box(g.uniform(-2, 2), g.uniform(-2, 2), g.uniform(2, 5), g.uniform(1, 3), g.uniform(-3, 3))
        worst = max(worst, abs(no.ref_overlap_f32(b1, b2) - no.bev_overlap(b1, b2)))
    assert worst < 0.1                                   # bounded by margin x side length
    iou = no.ref_iou_f32(np.stack([a, box(2, 0, 4, 2, 0)]), np.stack([a]))
    assert abs(iou[0, 0] - 1.0) < 1e-6 and abs(iou[1, 0] - 4.0 / 12.0) < 1e-6
    n = no.iou_normal_f32(np.stack([a, box(2, 0, 4, 2, 1.0)]), np.stack([a]))
    assert abs(n[0, 0] - 1.0) < 1e-6 and abs(n[1, 0] - 4.0 / 12.0) < 1e-6   # heading is ignored
