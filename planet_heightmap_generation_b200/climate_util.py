"""Host-side mirror of js/climate-util.js."""
from __future__ import annotations

from .engine import DeviceMesh


def smoothField(mesh: DeviceMesh, field, passes):
    """js/climate-util.js:5-25 — `passes` Laplacian sweeps, in place."""
    mesh._begin(field)
    mesh.lib.check(mesh.lib.dll.pb_smooth_field(mesh._mesh, mesh._ptr(field, "f32", mesh.numRegions, "field"), int(passes)))
