/* backing store for the memory-mapped FPGA_REG block that fpga.c pokes */
char u96_ref_regspace[1 << 16];

/* command words captured from issue_cmd() (see capture_cmd.h) */
static unsigned int u96_ref_cmd[1 << 16];
static int u96_ref_ncmd;
unsigned int u96_ref_capture(unsigned long v)
{
    if (u96_ref_ncmd < (int)(sizeof(u96_ref_cmd) / sizeof(u96_ref_cmd[0]))) u96_ref_cmd[u96_ref_ncmd] = (unsigned int)v;
    u96_ref_ncmd++;
    return (unsigned int)v;
}
void u96_ref_cmd_reset(void) { u96_ref_ncmd = 0; }
int u96_ref_cmd_count(void) { return u96_ref_ncmd; }
const unsigned int *u96_ref_cmd_data(void) { return u96_ref_cmd; }
