// Iteration cadence and tolerances of the trainers — same names and values as the reference's
// const.h:4-12 so that code written against it compiles unchanged.
#ifndef _CONST_H_
#define _CONST_H_

#define OBJ_ITER 1
#define DISP_ITER 50
#define SAVE_ITER 50
#define CHANCE_ITER 500
#define EPS 1e-5
#define GK_CSR_IS_VAL 1

#endif
