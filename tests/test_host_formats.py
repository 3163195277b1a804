"""CPU tests of the C++ host's file formats against independent implementations (scipy.io, a hand-rolled proto2
reader): MAT v5 both directions, protobuf text parsing of expopt / part_conf, joint loading + flip, .pbuf writing."""
import os
import struct
import subprocess

import numpy as np
import pytest
import scipy.io

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "partapp_b200", "psinfer_host_selftest")


@pytest.fixture(scope="module")
def tool(pslib):
    subprocess.run(["make", "-C", os.path.join(ROOT, "partapp_b200", "csrc", "host")], check=True, capture_output=True)
    return TOOL


def run(tool, *args):
    r = subprocess.run([tool] + list(args), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return r.stdout


@pytest.mark.parametrize("mode", ["compressed", "raw"])
def test_mat_writer_is_readable_by_scipy(tool, tmp_path, mode):
    f = str(tmp_path / "w.mat")
    run(tool, "mat-write", f, *([] if mode == "compressed" else ["raw"]))
    m = scipy.io.loadmat(f)
    # C dims are kept and element [i][j][k] stays element (i, j, k) (matlab_io.hpp:72-75)
    assert m["a"].dtype == np.float32 and np.array_equal(m["a"], (np.arange(24) * 0.5).reshape(2, 3, 4))
    assert m["b"].dtype == np.float64 and np.array_equal(m["b"], [[1.5, -2.5], [3.25, 4], [5, 6e10]])
    assert m["s"].shape == (1, 1) and m["s"][0, 0] == 7.5
    assert m["v"].shape == (3, 1)  # mat_save_std_vector writes n x 1 (matlab_io.cpp:158-166)


def test_mat_reader_reads_scipy_files(tool, tmp_path):
    f = str(tmp_path / "r.mat")
    cg = np.empty((1, 3), dtype=object)
    for r in range(3):
        cg[0, r] = (np.arange(6, dtype=np.float32).reshape(2, 3) + 10 * r)
    x = np.arange(24, dtype=np.float32).reshape(2, 3, 4) * 0.25
    scipy.io.savemat(f, {"x": x, "ints": np.array([[1.0, 2.0, 300.0]]), "cg": cg, "k": 3.5}, do_compression=True)
    out = run(tool, "mat-dump", f, "x").split("\n")
    assert out[0].startswith("name x class 7 dims 2 3 4")
    assert [float(v) for v in out[1:25]] == list(x.reshape(-1))          # C order
    out = run(tool, "mat-dump", f, "ints").split("\n")                    # stored as uint16 by savemat? any type is widened
    assert [float(v) for v in out[1:4]] == [1.0, 2.0, 300.0]
    out = run(tool, "mat-dump", f, "cg")
    assert "class 1 dims 1 3" in out.split("\n")[0]
    vals = [float(v) for v in out.split("\n") if v.strip() and not v.strip().startswith("name")]
    assert vals == [float(v) for r in range(3) for v in (np.arange(6) + 10 * r)]
    assert run(tool, "mat-dump", f, "k").split("\n")[1] == "3.5"


def _parse_hyps(b):
    def varint(i):
        v = s = 0
        while True:
            c = b[i]
            i += 1
            v |= (c & 0x7f) << s
            s += 7
            if not c & 0x80:
                return v, i
    i, out = 0, []
    while i < len(b):
        key, i = varint(i)
        assert key == (1 << 3 | 2)
        n, i = varint(i)
        end, h = i + n, {}
        while i < end:
            k, i = varint(i)
            if k & 7 == 5:
                h[k >> 3] = struct.unpack("<f", b[i:i + 4])[0]
                i += 4
            else:
                h[k >> 3], i = varint(i)
        out.append(h)
    return out


def test_hypothesis_list_wire_format(tool, tmp_path):
    f = str(tmp_path / "h.pbuf")
    run(tool, "pbuf-write", f)
    hyps = _parse_hyps(open(f, "rb").read())
    assert len(hyps) == 3
    for i, h in enumerate(hyps):   # HypothesisList.proto: x=1, y=2, scale=3, score=4 (float), flip=5 (bool)
        assert h[1] == 10.0 + i and h[2] == 20.0 + i and abs(h[3] - (1 + 0.1 * i)) < 1e-6 and h[4] == -3.5 * i
        assert h[5] == (1 if i == 1 else 0)
    try:  # cross-check with protobuf-python if it is importable
        from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
        fd = descriptor_pb2.FileDescriptorProto(name="HypothesisList.proto", syntax="proto2")
        m = fd.message_type.add(name="HypothesisList")
        oh = m.nested_type.add(name="ObjectHypothesis")
        for n_, num in (("x", 1), ("y", 2), ("scale", 3), ("score", 4)):
            oh.field.add(name=n_, number=num, type=2, label=1)
        oh.field.add(name="flip", number=5, type=8, label=1)
        m.field.add(name="hyp", number=1, type=11, label=3, type_name=".HypothesisList.ObjectHypothesis")
        pool = descriptor_pool.DescriptorPool()
        pool.Add(fd)
        cls = message_factory.GetMessageClass(pool.FindMessageTypeByName("HypothesisList"))
        msg = cls()
        msg.ParseFromString(open(f, "rb").read())
        assert len(msg.hyp) == 3 and msg.hyp[1].flip and msg.hyp[2].score == -7.0
    except ImportError:
        pass


def test_expopt_part_conf_joints(tool, tmp_path):
    from tests.make_experiment import make
    info = make(str(tmp_path / "exp"), num_images=2)
    out = run(tool, "expopt-dump", info["expopt"], "0")
    lines = dict(l.split(" ", 1) for l in out.strip().split("\n") if not l.startswith(("joint", "image", "J ")))
    assert lines["log_subdir"] == "exp-synth"                        # defaults to the expopt basename (partapp.cpp:313-318)
    assert lines["scoregrid_dir"].endswith("log_dir/exp-synth/test_scoregrid")
    assert lines["spatial_dir"].endswith("log_dir/exp-synth/spatial")
    assert lines["rot"].startswith("8 -180 180 scale 1 1 1")
    assert lines["parts"] == "4 root %d joints 3" % info["root_idx"]
    assert lines["bbox_offset"] == "3.7 -2.2"
    imgs = [l.split() for l in out.split("\n") if l.startswith("image")]
    assert len(imgs) == 2 and imgs[0][2:] == ["40", "48"]
    js = [l.split()[1:] for l in out.split("\n") if l.startswith("J ")]
    for row, j in zip(js, info["joints"]):
        vals = [float(v) for v in row]
        assert vals[:3] == [2, j.child_idx, j.parent_idx]            # 0-based after loadJoints (aux.cpp:123-124)
        assert vals[3:5] == list(j.offset_c) and vals[5:7] == list(j.offset_p)
        assert vals[7:11] == list(np.asarray(j.C).reshape(4)) and vals[11:] == [j.rot_mean, j.rot_sigma]
    # flip (aux.cpp:102-119)
    out = run(tool, "expopt-dump", info["expopt"], "1")
    js = [l.split()[1:] for l in out.split("\n") if l.startswith("J ")]
    for row, j in zip(js, info["joints"]):
        vals = [float(v) for v in row]
        assert vals[3] == -j.offset_c[0] and vals[4] == j.offset_c[1] and vals[11] == -j.rot_mean
        assert vals[8] == -np.asarray(j.C)[0, 1]


@pytest.mark.parametrize("npos", [1, 2, 4])
def test_use_gt_torso_reads_the_annotated_root_position(tool, tmp_path, npos):
    """ExpParam.use_gt_torso (objectdetect_icps.cpp:292-300): the root position prior is centred on part_pos of
    get_part_bbox(first annotated rectangle, root part), truncated to int -- checked against parteval.get_part_bbox, the
    Python mirror that tests/test_eval_vs_ref.py pins to the reference's own partdef.cpp."""
    from partapp_b200 import parteval as pe
    from tests.make_experiment import make
    info = make(str(tmp_path / "exp"), num_images=3, extra_expopt="use_gt_torso: true\n")
    root = info["root_idx"]
    ids = [11, 12, 13, 14][:npos]
    pts = [[(17, 9), (30, 41), (3, 22), (25, 25)], [(5, 5), (8, 12), (39, 1), (20, 47)], [(1, 2), (2, 1), (0, 0), (7, 7)]]
    conf = ""
    for p in range(info["P"]):
        conf += "part {\n  part_id: %d\n  is_detect: true\n  is_root: %s\n" % (p + 1, "true" if p == root else "false")
        if p == root:
            conf += "".join("  part_pos: %d\n" % i for i in ids) + "  part_x_axis_from: 11\n  part_x_axis_to: 11\n"
        conf += "}\n"
    edges = [(j.child_idx, j.parent_idx) for j in info["joints"]]
    conf += "".join('joint {\n  child_idx: %d\n  parent_idx: %d\n  type: "RotGaussian"\n}\n' % (c + 1, q + 1) for c, q in edges)
    (tmp_path / "exp" / "part_conf.txt").write_text(conf)
    al = "<annotationlist>\n"
    for i in range(3):
        point = lambda k, xy: "<point><id>%d</id><x>%d</x><y>%d</y></point>" % (k, xy[0], xy[1])
        first = "".join(point(11 + k, pts[i][k]) for k in (2, 0, 3, 1)) + point(11, (99, 99))   # a duplicate id: first wins
        second = "".join(point(11 + k, (50, 50)) for k in range(4))                                # another person: ignored
        al += ("<annotation><image><name>images/im%04d.png</name></image><annorect><x1>1</x1><y1>1</y1><x2>9</x2><y2>9</y2>"
               "<annopoints>%s</annopoints></annorect><annorect><annopoints>%s</annopoints></annorect></annotation>\n"
               % (i, first, second))
    (tmp_path / "exp" / "test.al").write_text(al + "</annotationlist>\n")
    out = run(tool, "expopt-dump", info["expopt"])
    got = [l.split()[1:] for l in out.split("\n") if l.startswith("gt_torso")]
    assert len(got) == 3
    pd = pe.PartDef(root + 1, ids, [11], [11])      # coinciding axis points: the axis is invalid, part_pos is still used
    for i in range(3):
        rect = pe.load_annolist(str(tmp_path / "exp" / "test.al"))[i].rects[0]
        if npos < 3:
            want = np.mean([rect.points[k] for k in ids], axis=0) if npos == 1 else \
                (np.sum([rect.points[k] for k in ids], axis=0) * (1.0 / npos))
        else:
            xs, ys = [rect.points[k][0] for k in ids], [rect.points[k][1] for k in ids]
            want = np.array([0.5 * (min(xs) + max(xs)), 0.5 * (min(ys) + max(ys))])
        assert [float(v) for v in got[i][1:]] == [float(int(want[0])), float(int(want[1]))], (i, got[i], want)
    # the same numbers from the pinned Python mirror where its axis is valid
    pd2 = pe.PartDef(root + 1, ids, [11], [12]) if npos >= 2 else None
    if pd2:
        rect = pe.load_annolist(str(tmp_path / "exp" / "test.al"))[0].rects[0]
        bb = pe.get_part_bbox(rect, pd2, 1.0)
        assert [float(v) for v in got[0][1:]] == [float(int(bb.part_pos[0])), float(int(bb.part_pos[1]))]
