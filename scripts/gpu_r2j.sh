python scripts/e2e_probe.py slots3
TRRE_B200_LIB=build/exp/libtrre_b200_slots4.so python scripts/e2e_probe.py slots4 | grep scan_into
TRRE_B200_LIB=build/exp/libtrre_b200_slots6.so python scripts/e2e_probe.py slots6 | grep scan_into
