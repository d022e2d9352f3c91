/* backing store for the memory-mapped FPGA_REG block that fpga.c pokes */
char u96_ref_regspace[1 << 16];
