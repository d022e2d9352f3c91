/* stub of the Xilinx BSP header: gives fpga.c a host-side register space */
#ifndef XPARAMETERS_H
#define XPARAMETERS_H
extern char u96_ref_regspace[];
#define XPAR_DVP_0_BASEADDR u96_ref_regspace
#endif
