/* cudadefs.h — kept so sources that include the reference's cudadefs.h keep compiling.
 * The reference hard-codes 64-thread blocks and a Maxwell target here (cudadefs.h:1-9); the B200 engine picks its
 * launch shapes per kernel, so only the index macro survives. */
#ifndef B2NN_CUDADEFS_H
#define B2NN_CUDADEFS_H

#ifndef IDX2C
#define IDX2C(i, j, ld) ((j) * (ld) + (i)) /* column-major: i = row (sample), j = column, ld = number of rows */
#endif

#endif
