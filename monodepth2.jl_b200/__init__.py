"""B200-native view-synthesis loss path of Monodepth2.jl (host-side mirror of the reference API)."""
