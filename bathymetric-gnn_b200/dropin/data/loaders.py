"""Drop-in for data/loaders.py.  The in-memory container is the path's input type; the file formats are not on the
path: `BathymetricLoader.load` / `BathymetricWriter.save` handle .npz themselves and hand every other format to the
reference's GDAL loaders (reference checkout + `osgeo` required, exactly as for the reference itself)."""
from bathymetric_gnn_b200.tiling import BathymetricGrid  # noqa: F401
from bathymetric_gnn_b200.pipeline import NpzLoader, NpzWriter, FormatDispatchLoader as BathymetricLoader, FormatDispatchWriter as BathymetricWriter  # noqa: F401
