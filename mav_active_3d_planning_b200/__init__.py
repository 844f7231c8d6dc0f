"""B200-native view-gain evaluation for mav_active_3d_planning (one hot path)."""
