#!/usr/bin/env python
"""Extracts the reference's numeric literals (the only golden vectors it holds) into
tests/golden/reference_constants.json.  Run in the dev container, where /root/reference exists:

    python tests/golden/make_golden.py

Sources: sdf/cube.sdf (robot), launch/cdpr_gazebo.launch (controller gains), src/sinevelocitytest.cpp and the
two square-wave drivers (command signals), include/cdpr_gazebo/CdprGazeboPlugin.h (cable count)."""
import json
import os
import re
import xml.etree.ElementTree as ET

REF = "/root/reference/src/cdpr_gazebo"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_constants.json")


def floats(text):
    return [float(x) for x in text.split()]


def main():
    sdf = ET.parse(os.path.join(REF, "sdf/cube.sdf")).getroot()
    model = sdf.find("model")
    links = {l.get("name"): l for l in model.findall("link")}
    joints = {j.get("name"): j for j in model.findall("joint")}
    g = {"source": "balazs-bamer/cdpr-simulation: sdf/cube.sdf, launch/cdpr_gazebo.launch, src/*test.cpp"}
    plat = links["platform"]
    g["platform_pose"] = floats(plat.find("pose").text)
    g["platform_mass"] = float(plat.find("inertial/mass").text)
    g["platform_inertia"] = [float(plat.find("inertial/inertia/" + k).text) for k in ("ixx", "iyy", "izz", "ixy", "ixz", "iyz")]
    g["frame_mass"] = float(links["frame"].find("inertial/mass").text)
    cables = []
    i = 0
    while f"cable{i}" in joints:
        j = joints[f"cable{i}"]
        ax = j.find("axis")
        cables.append({
            "cable_link_pose": floats(links[f"cable{i}"].find("pose").text),
            "frame_anchor_link_pose": floats(links[f"virt_X{i}"].find("pose").text),
            "platform_anchor_link_pose": floats(links[f"virt_Xpf{i}"].find("pose").text),
            "prismatic_axis": floats(ax.find("xyz").text),
            "lower": float(ax.find("limit/lower").text), "upper": float(ax.find("limit/upper").text),
            "effort": float(ax.find("limit/effort").text), "velocity": float(ax.find("limit/velocity").text),
            "damping": float(ax.find("dynamics/damping").text),
            "parent": j.find("parent").text, "child": j.find("child").text,
            "passive_damping": float(joints[f"rev_X{i}"].find("axis/dynamics/damping").text),
            "leg_link_mass": float(links[f"virt_X{i}"].find("inertial/mass").text),
            # the five links of the leg and its five passive revolute joints
            "leg_links": {nm: {"mass": float(links[f"{nm}{i}"].find("inertial/mass").text),
                               "inertia": [float(links[f"{nm}{i}"].find("inertial/inertia/" + k).text) for k in ("ixx", "iyy", "izz", "ixy", "ixz", "iyz")],
                               "pose": floats(links[f"{nm}{i}"].find("pose").text)}
                          for nm in ("cable", "virt_X", "virt_Y", "virt_Xpf", "virt_Ypf")},
            "leg_joints": {nm: {"axis": floats(joints[f"{nm}{i}"].find("axis/xyz").text), "parent": joints[f"{nm}{i}"].find("parent").text,
                                "child": joints[f"{nm}{i}"].find("child").text, "damping": float(joints[f"{nm}{i}"].find("axis/dynamics/damping").text)}
                           for nm in ("rev_X", "rev_Y", "rev_Xpf", "rev_Ypf", "rev_Zpf")},
        })
        i += 1
    g["cables"] = cables
    g["n_links"] = len(links)
    g["n_joints"] = len(joints)
    g["sdf_version"] = sdf.get("version")
    launch = ET.parse(os.path.join(REF, "launch/cdpr_gazebo.launch")).getroot()
    g["launch_params"] = {p.get("name"): float(p.get("value")) for p in launch.iter("param")}
    hdr = open(os.path.join(REF, "include/cdpr_gazebo/CdprGazeboPlugin.h")).read()
    g["wire_count"] = int(re.search(r"cWireCount\s*=\s*(\d+)u", hdr).group(1))
    drivers = {}
    for name in ("sinevelocitytest", "squarevelocitytest", "squarepositiontest"):
        src = open(os.path.join(REF, f"src/{name}.cpp")).read()
        drivers[name] = {m.group(1): float(m.group(2)) for m in re.finditer(r"const double (c\w+)\s*=\s*([0-9.eE+-]+);", src)}
    g["drivers"] = drivers
    import yaml
    g["cube_yaml"] = yaml.safe_load(open(os.path.join(REF, "sdf/cube.yaml")))
    json.dump(g, open(OUT, "w"), indent=1, sort_keys=True)
    print("wrote", OUT, "cables:", len(cables))


if __name__ == "__main__":
    main()
