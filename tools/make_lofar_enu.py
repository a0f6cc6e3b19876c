#!/usr/bin/env python
"""Derive ionotomo_b200/data/lofar_hba_enu_km.txt from the reference's station list
(astro/arrays/lofar.hba.antenna.cfg, ITRS X Y Z in metres): local East-North-Up
coordinates in km about the array centroid (WGS84 geodetic lat/lon of the centroid),
the frame the benchmark's synthetic rays are cast in.  Build-container
only; the derived table is what ships."""
import numpy as np

SRC = "/root/reference/src/ionotomo/astro/arrays/lofar.hba.antenna.cfg"
DST = "ionotomo_b200/data/lofar_hba_enu_km.txt"

names, xyz = [], []
for line in open(SRC):
    line = line.strip()
    if not line or line.startswith("#"):
        continue
    p = line.split()
    xyz.append([float(p[0]), float(p[1]), float(p[2])])
    names.append(p[4])
xyz = np.array(xyz)
c = xyz.mean(0)
# WGS84 geodetic latitude/longitude of the centroid (Bowring's iteration)
a, f = 6378137.0, 1 / 298.257223563
e2 = f * (2 - f)
lon = np.arctan2(c[1], c[0])
pxy = np.hypot(c[0], c[1])
lat = np.arctan2(c[2], pxy * (1 - e2))
for _ in range(10):
    Nn = a / np.sqrt(1 - e2 * np.sin(lat) ** 2)
    lat = np.arctan2(c[2] + e2 * Nn * np.sin(lat), pxy)
east = np.array([-np.sin(lon), np.cos(lon), 0.])
north = np.array([-np.sin(lat) * np.cos(lon), -np.sin(lat) * np.sin(lon), np.cos(lat)])
up = np.array([np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)])
R = np.stack([east, north, up])
enu = (xyz - c) @ R.T / 1000.0
with open(DST, "w") as fh:
    fh.write("# LOFAR HBA stations, local ENU about the array centroid, km\n")
    fh.write("# centroid geodetic lat %.6f deg, lon %.6f deg (WGS84)\n" % (np.degrees(lat), np.degrees(lon)))
    fh.write("# east_km north_km up_km station\n")
    for n, r in zip(names, enu):
        fh.write("%.6f %.6f %.6f %s\n" % (r[0], r[1], r[2], n))
print(len(names), "stations; E %.1f..%.1f N %.1f..%.1f U %.2f..%.2f km" % (
    enu[:, 0].min(), enu[:, 0].max(), enu[:, 1].min(), enu[:, 1].max(), enu[:, 2].min(), enu[:, 2].max()))
