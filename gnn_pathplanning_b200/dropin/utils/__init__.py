# Drop-in namespace: resolves the reference's `utils.graphUtils.graphML` import to gnn_pathplanning_b200.
