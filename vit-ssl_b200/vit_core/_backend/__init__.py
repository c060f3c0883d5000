"""Host-side binding layer between the `vit_core` modules and the sm_100a C-ABI library."""
