"""Contact variants of the corpus YAMLs for the contact-dynamics tests (test infrastructure): written into a private
YAML tree under tmp_path, selected with EAGLE_MPC_YAML_DIR."""
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
YAML_ROOT = os.path.join(ROOT, "yaml")


def six_d(tmp_path):
    """monkey_bar.yaml with its ContactModel3D turned into a ContactModel6D"""
    root = tmp_path / "yaml"
    shutil.copytree(os.path.join(YAML_ROOT, "hexacopter370_flying_arm_3"), root / "hexacopter370_flying_arm_3")
    src = (root / "hexacopter370_flying_arm_3" / "trajectories" / "monkey_bar.yaml").read_text()
    assert 'type: "ContactModel3D"' in src
    src = src.replace('type: "ContactModel3D"', 'type: "ContactModel6D"\n          orientation: [0, 0, 0, 1]')
    (root / "hexacopter370_flying_arm_3" / "trajectories" / "monkey_bar_6d.yaml").write_text(src)
    return str(root), "hexacopter370_flying_arm_3/trajectories/monkey_bar_6d.yaml"


def hextilt_push(tmp_path):
    """push_slide.yaml (hextilt + 5-joint arm, the widest platform) with its gripper in contact and a friction cone on
    the contact force: the contact path on Dim<5, 6>"""
    root = tmp_path / "yaml"
    shutil.copytree(os.path.join(YAML_ROOT, "hextilt_flying_arm_5"), root / "hextilt_flying_arm_5")
    shutil.copytree(os.path.join(YAML_ROOT, "hextilt"), root / "hextilt", dirs_exist_ok=True)
    p = root / "hextilt_flying_arm_5" / "trajectories" / "push_slide.yaml"
    src = p.read_text().rstrip("\n")
    lines = src.split("\n")
    # indentation of the stage's "costs:" key
    ci = next(i for i, l in enumerate(lines) if l.strip() == "costs:")
    ind = lines[ci][:len(lines[ci]) - len(lines[ci].lstrip())]
    item = ind + "  "
    cone = [item + '- name: "friction_cone"', item + '  type: "CostModelContactFrictionCone"', item + "  weight: 10", item + "  n_surf: [0.2, 0.1, 1]",
            item + "  mu: 0.6", item + '  link_name: "flying_arm_5__gripper"']
    contact = [ind + "contacts:", item + '- name: "end_effector"', item + '  type: "ContactModel3D"', item + '  link_name: "flying_arm_5__gripper"',
               item + "  position: [0, 0, 0]", item + "  gains: [0, 0]"]
    out = lines[:ci + 1] + cone + lines[ci + 1:] + contact
    (root / "hextilt_flying_arm_5" / "trajectories" / "push_contact.yaml").write_text("\n".join(out) + "\n")
    return str(root), "hextilt_flying_arm_5/trajectories/push_contact.yaml"
