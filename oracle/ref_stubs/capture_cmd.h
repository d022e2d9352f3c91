/* Force-included (-include) in front of the reference's UNMODIFIED fpga.c: records every word issue_cmd() writes to the
 * memory-mapped rectifier command port (fpga.c:505-547 -> FPGA_REG_RECT.CmdWritePort, fpga.h:183), so that the command
 * stream the reference generates can be compared with the reference's own shipped dump src/dvp/sim/cmd.dat.
 * The struct is declared first (fpga.h is include-guarded); afterwards every statement `fpga->rect.CmdWritePort = val;`
 * expands to `fpga->rect.CmdWritePort = u96_ref_capture(val); fpga->rect.CmdWritePort = val;`. */
#ifndef U96_REF_CAPTURE_CMD_H
#define U96_REF_CAPTURE_CMD_H
#include "fpga.h"
unsigned int u96_ref_capture(unsigned long v);
#define CmdWritePort CmdWritePort = u96_ref_capture(val); fpga->rect.CmdWritePort
#endif
