/* GlobalGate -- scripts of the globalCondition test model (authored for this repository, stock model API). */
#ifndef _FUNCTIONS_H_
#define _FUNCTIONS_H_

#include "header.h"

__FLAME_GPU_FUNC__ int raise(xmachine_memory_Node* agent)
{
	agent->level += 1 + (agent->id % 3);      /* nodes climb at three different speeds */
	return 0;
}

__FLAME_GPU_FUNC__ int release(xmachine_memory_Node* agent)
{
	agent->level = agent->id % 2;
	agent->rounds++;
	return 0;
}

#endif
