"""Drop-in for data/vr_bag.py."""
from bathymetric_gnn_b200.vr_bag import (NODATA, RefinementGrid, RefinementTable, VRBagHandler, VRBagWriter,  # noqa: F401
                                         SRBagHandler, SRBagWriter, detect_bag_type, process_vr_bag_native)
from bathymetric_gnn_b200.native_vr import SidecarBuilder  # noqa: F401
