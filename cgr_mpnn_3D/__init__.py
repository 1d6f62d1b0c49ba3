"""Drop-in import path of the reference package (``cgr_mpnn_3D.models.GNN``), backed by cgr_mpnn_3d_b200."""
