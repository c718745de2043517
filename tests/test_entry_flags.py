"""The entry points keep the reference's command-line interface: every option string of /root/reference's
Train_Stage1_K.py / Train_Stage2_K.py / Test_KITTI.py (extracted into tests/golden/entry_flags.json by
tests/golden/make_flags.py) is accepted, training hyper-parameter defaults are the reference's, and a
reference-style command line parses.  Data-loading / dump options are accepted and unused (out of scope)."""
import importlib
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REF = json.load(open(os.path.join(HERE, "golden", "entry_flags.json")))

# defaults that are machine-specific in the reference (Windows paths, time stamps, a resume epoch) or deliberately
# different here (Test_KITTI batch: the reference's loader is batch 1, ours shards 8 images per GPU)
SKIP_DEFAULT = {"--data", "--pretrained", "--fix_model", "--start-epoch", "--time_stamp", "--details", "--gpu_no",
                "--dataset"}
OURS_DIFFER = {("Test_KITTI", "--batch_size")}


@pytest.mark.parametrize("name", sorted(REF))
def test_reference_flags_are_accepted(name):
    mod = importlib.import_module(name)
    ours = {o: a for a in mod.parser._actions for o in a.option_strings}
    for f in REF[name]:
        for o in f["options"]:
            assert o in ours, (name, o)
        long = next(o for o in f["options"] if o.startswith("--"))
        if "default" in f and f["default"] != "<expr>" and long not in SKIP_DEFAULT and (name, long) not in OURS_DIFFER:
            assert ours[long].default == f["default"], (name, long, ours[long].default, f["default"])


def test_reference_style_command_lines_parse():
    s1 = importlib.import_module("Train_Stage1_K").parser.parse_args(
        "-d /data -n0 Kitti -train_split eigen_train_split -vdn Kitti2015 -maxd 300 -mind 2 -b 8 -ch 192 -cw 640 "
        "-perc 0.01 -smooth 0.0008 --lr 0.0001 --milestones 30 40 --epochs 50 -w 4 -tbs 1".split())
    assert s1.batch_size == 8 and s1.milestones == [30, 40] and s1.a_p == 0.01
    s2 = importlib.import_module("Train_Stage2_K").parser.parse_args("-mirror_loss 1 -b 4 --fix_model ckpt.pth.tar".split())
    assert s2.a_mr == 1 and s2.lr == 0.00005 and s2.epochs == 20
    t = importlib.import_module("Test_KITTI").parser
    a = t.parse_args("-tn Kitti_eigen_test_improved -fpp True -mspp False -save False -m FAL_netB".split())
    assert a.f_post_process is True and a.ms_post_process is False
    b = t.parse_args([])
    assert b.f_post_process is False and b.ms_post_process is True      # the reference's shipped default: multi-scale PP
    c = t.parse_args(["-fpp"])
    assert c.f_post_process is True
