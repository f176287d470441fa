extern MAC_xface *mac_xface;
