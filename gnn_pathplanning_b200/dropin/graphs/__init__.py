# Drop-in namespace: resolves the reference's `graphs.*` imports to gnn_pathplanning_b200.
