"""Layer tables of the three auto-label models (the reference's module layout, which fixes the
``state_dict`` contract of SURVEY.md section 8b).

Reference: tools/static_model.py:241-269 (seg net), :298-318 (static box head),
tools/dynamic_model.py:214-232 (PointEmbedding), :251-269 (BoxEmbedding), :288-298 (dynamic head).
Every entry is (layer name, bn name or None, in_features, out_features, kind) with kind in
{"conv", "fc"}; conv weights are stored (out, in, 1), fc weights (out, in).
"""

NUM_HEADING_BIN = 12
NUM_SIZE_CLUSTER = 3
NUM_OBJECT_POINT = 512
NUM_POINT_STATIC = 4096
NUM_POINT_DYNAMIC = 1024
NUM_FRAME = 5
NUM_BOX_STEPS = 101
HEAD_WIDTH = 3 + NUM_HEADING_BIN * 2 + NUM_SIZE_CLUSTER * 4   # 39
MEAN_SIZE_ARR = ((4.8, 1.8, 1.5), (10.0, 2.6, 3.2), (2.0, 1.0, 1.6))
BN_EPS = 1e-5


def seg_layers(n_channel):
    return [
        ("conv1", "bn1", n_channel, 64, "conv"), ("conv2", "bn2", 64, 64, "conv"),
        ("conv3", "bn3", 64, 64, "conv"), ("conv4", "bn4", 64, 128, "conv"),
        ("conv5", "bn5", 128, 1024, "conv"),
        ("dconv1", "dbn1", 1088, 512, "conv"), ("dconv2", "dbn2", 512, 256, "conv"),
        ("dconv3", "dbn3", 256, 128, "conv"), ("dconv4", "dbn4", 128, 128, "conv"),
        ("dconv5", None, 128, 2, "conv"),
    ]


def static_est_layers():
    return [
        ("conv1", "bn1", 3, 128, "conv"), ("conv2", "bn2", 128, 128, "conv"),
        ("conv3", "bn3", 128, 256, "conv"), ("conv4", "bn4", 256, 512, "conv"),
        ("fc1", "fcbn1", 512, 512, "fc"), ("fc2", "fcbn2", 512, 256, "fc"),
        ("fc3", None, 256, HEAD_WIDTH, "fc"),
    ]


def point_emb_layers():
    return [
        ("conv1", "bn1", 4, 64, "conv"), ("conv2", "bn2", 64, 128, "conv"),
        ("conv3", "bn3", 128, 256, "conv"), ("conv4", "bn4", 256, 512, "conv"),
        ("fc1", "fcbn1", 512, 512, "fc"), ("fc2", "fcbn2", 512, 256, "fc"),
    ]


def box_emb_layers():
    return [
        ("conv1", "bn1", 8, 64, "conv"), ("conv2", "bn2", 64, 64, "conv"),
        ("conv3", "bn3", 64, 128, "conv"), ("conv4", "bn4", 128, 512, "conv"),
        ("fc1", "fcbn1", 512, 128, "fc"), ("fc2", "fcbn2", 128, 128, "fc"),
    ]


def dynamic_est_layers(n_classes=3):
    return [
        ("fc1", "fcbn1", 256 + 128, 128, "fc"), ("fc2", "fcbn2", 128, 128, "fc"),
        ("fc3", None, 128, n_classes + NUM_HEADING_BIN * 2 + NUM_SIZE_CLUSTER * 4, "fc"),
    ]


def model_blocks(kind, n_channel=None):
    """Ordered (submodule name, layer table) pairs of a model kind."""
    if kind == "static_one":
        return [("ins_seg", seg_layers(n_channel or 3)), ("box_est", static_est_layers())]
    if kind == "static_two":
        return [("ins_seg", seg_layers(n_channel or 3)), ("box_est_one", static_est_layers()),
                ("box_est_two", static_est_layers())]
    if kind == "dynamic":
        return [("ins_seg", seg_layers(n_channel or 4)), ("point_emb", point_emb_layers()),
                ("box_emb", box_emb_layers()), ("box_est", dynamic_est_layers())]
    raise ValueError("unknown model kind %r" % (kind,))


def flops_per_object(kind, n_points):
    """Factored algorithmic FLOPs per object: 1 MAC = 2 FLOP and the
    1024-wide global-feature half of ins_seg.dconv1 is counted once per object, not per point."""
    macs = 1024 * 512
    for name, table in model_blocks(kind):
        for lname, _bn, cin, cout, lkind in table:
            if name == "ins_seg":
                cin_eff = 64 if lname == "dconv1" else cin
                macs += cin_eff * cout * n_points
            elif lkind == "conv":
                m = {"box_est": NUM_OBJECT_POINT, "box_est_one": NUM_OBJECT_POINT, "box_est_two": NUM_OBJECT_POINT,
                     "point_emb": NUM_FRAME * NUM_OBJECT_POINT, "box_emb": NUM_BOX_STEPS}[name]
                macs += cin * cout * m
            else:
                macs += cin * cout
    return 2.0 * macs
