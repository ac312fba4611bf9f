/*
 * IdSort -- scripts for the id-generation and agent-sort fixture.
 *
 * Authored for this repository with the stock model API only:
 *   take_ticket   every agent draws a fresh id on the DEVICE (generate_Item_id, an atomic counter);
 *   reorder       a step function sorts the list by (ticket % 7) with sort_Items_default and its own
 *                 key/value kernel -- many equal keys, so the order among them shows whether the sort is stable;
 *   add_one       a step function creates one agent on the HOST whose id comes from the same counter.
 */
#ifndef _FUNCTIONS_H_
#define _FUNCTIONS_H_

#include "header.h"

__FLAME_GPU_FUNC__ int take_ticket(xmachine_memory_Item* agent)
{
	agent->ticket = generate_Item_id();
	return 0;
}

#ifdef FGPU_B200
__global__ void ticket_keys(unsigned int* keys, unsigned int* values, xmachine_memory_Item_list* agents)
{
	int index = (blockIdx.x * blockDim.x) + threadIdx.x;
	keys[index] = (unsigned int)(agents->ticket[index] % 7);
	values[index] = index;
}
#endif

__FLAME_GPU_STEP_FUNC__ void reorder()
{
#ifdef FGPU_B200
	sort_Items_default(&ticket_keys);
#endif
}

__FLAME_GPU_STEP_FUNC__ void add_one()
{
	xmachine_memory_Item* item = h_allocate_agent_Item();
	item->id = generate_Item_id();
	item->x = -1.0f;
	item->ticket = -1;
	h_add_agent_Item_default(item);
	h_free_agent_Item(&item);
}

#endif
