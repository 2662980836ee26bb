#!/usr/bin/env python
"""Writes the PLY fixtures of tests/golden/: files in the exact layout pcl::PLYWriter produces for a cloud with fields
x,y,z,rgba (what the reference's cwipc_write emits, src/cwipc_util.cpp:466-497): `comment PCL generated`, vertex
properties float x,y,z + uchar red,green,blue,alpha (the tile number travels in alpha), then PCL's one-row `camera`
element, ASCII and binary_little_endian; plus the two packed-colour layouts older PCL writers used (`uint rgba`,
`float rgb` = the packed word's bits carried in a float).  Nothing here comes from our own writer, so reading these
files pins cwipc_read against the format and not against itself.

    python tests/golden/make_ply_fixtures.py
"""
import os
import struct

HERE = os.path.dirname(os.path.abspath(__file__))

# x, y, z, r, g, b, tile
POINTS = [
    (0.0, 0.0, 0.0, 0, 0, 0, 0),
    (0.125, -1.5, 2.25, 255, 128, 1, 1),
    (-0.0012345679, 1.9999999, 0.33333334, 17, 34, 51, 2),
    (1e-7, 123456.79, -3.4e10, 1, 2, 3, 4),
    (0.3, 1.0, -0.3, 200, 100, 50, 255),
]

CAMERA_PROPS = ["view_px", "view_py", "view_pz", "x_axisx", "x_axisy", "x_axisz", "y_axisx", "y_axisy", "y_axisz", "z_axisx", "z_axisy", "z_axisz",
                "focal", "scalex", "scaley", "centerx", "centery"]


def pcl_header(fmt, n):
    lines = ["ply", f"format {fmt} 1.0", "comment PCL generated", f"element vertex {n}",
             "property float x", "property float y", "property float z",
             "property uchar red", "property uchar green", "property uchar blue", "property uchar alpha",
             "element camera 1"]
    lines += [f"property float {p}" for p in CAMERA_PROPS]
    lines += ["property int viewportx", "property int viewporty", "property float k1", "property float k2", "end_header"]
    return "\n".join(lines) + "\n"


def f32(v):
    return struct.unpack("<f", struct.pack("<f", v))[0]


def main():
    n = len(POINTS)
    cam_floats = [0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 0, 0]
    with open(os.path.join(HERE, "pcl_ascii.ply"), "w") as f:
        f.write(pcl_header("ascii", n))
        for x, y, z, r, g, b, t in POINTS:
            f.write(f"{f32(x)!r} {f32(y)!r} {f32(z)!r} {r} {g} {b} {t}\n")
        f.write(" ".join(str(v) for v in cam_floats) + f" {n} 1 0 0\n")
    with open(os.path.join(HERE, "pcl_binary.ply"), "wb") as f:
        f.write(pcl_header("binary_little_endian", n).encode())
        for x, y, z, r, g, b, t in POINTS:
            f.write(struct.pack("<fffBBBB", x, y, z, r, g, b, t))
        f.write(struct.pack("<17f", *cam_floats) + struct.pack("<ii", n, 1) + struct.pack("<ff", 0, 0))
    with open(os.path.join(HERE, "packed_uint_rgba.ply"), "w") as f:
        f.write(f"ply\nformat ascii 1.0\ncomment PCL generated\nelement vertex {n}\nproperty float x\nproperty float y\nproperty float z\nproperty uint rgba\nend_header\n")
        for x, y, z, r, g, b, t in POINTS:
            f.write(f"{f32(x)!r} {f32(y)!r} {f32(z)!r} {(t << 24) | (r << 16) | (g << 8) | b}\n")
    with open(os.path.join(HERE, "packed_float_rgb.ply"), "wb") as f:
        f.write(f"ply\nformat binary_little_endian 1.0\ncomment PCL generated\nelement vertex {n}\nproperty float x\nproperty float y\nproperty float z\nproperty float rgb\nend_header\n".encode())
        for x, y, z, r, g, b, t in POINTS:
            f.write(struct.pack("<fffI", x, y, z, (r << 16) | (g << 8) | b))   # alpha absent in the legacy layout


if __name__ == "__main__":
    main()
