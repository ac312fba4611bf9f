/*
 * BirthDeath -- agent scripts for fixture C5' (SURVEY.md 8d).
 *
 * No BASELINE-config model of the reference removes agents from a list on the
 * spatially partitioned path (Keratinocyte "death" is a type change), so this
 * authored model is the fixture that reaches death compaction
 * (gpu:reallocate = true, simulation.xslt:1763-1792) together with agent birth
 * into the function's own next state (newborn-first ordering,
 * simulation.xslt:1794-1829 then :1955) and rand48.  Stock model API only.
 *
 * Layers:  output_location -> live
 */
#ifndef _FUNCTIONS_H_
#define _FUNCTIONS_H_

#include "header.h"

__FLAME_GPU_FUNC__ int output_location(xmachine_memory_Cell* agent, xmachine_message_location_list* location_messages)
{
	add_location_message(location_messages, agent->id, agent->x, agent->y, agent->z);
	return 0;
}

__FLAME_GPU_FUNC__ int live(xmachine_memory_Cell* agent, xmachine_memory_Cell_list* Cell_agents, xmachine_message_location_list* location_messages, xmachine_message_location_PBM* partition_matrix, RNG_rand48* rand48)
{
	/* crowding: number of messages in the 27-cell neighbourhood (self included) */
	int crowd = 0;
	xmachine_message_location* msg = get_first_location_message(location_messages, partition_matrix, agent->x, agent->y, agent->z);
	while (msg)
	{
		crowd++;
		msg = get_next_location_message(msg, location_messages, partition_matrix);
	}

	agent->age++;
	agent->crowd = crowd;

	const float u = rnd<CONTINUOUS>(rand48);
	if (u < DEATH_P)
		return 1;                 /* removed from the list by the death compaction */
	if (u > 1.0f - BIRTH_P)
	{
		const float jx = rnd<CONTINUOUS>(rand48) - 0.5f;
		const float jy = rnd<CONTINUOUS>(rand48) - 0.5f;
		const float jz = rnd<CONTINUOUS>(rand48) - 0.5f;
		add_Cell_agent(Cell_agents, agent->id ^ 0x40000000, agent->x + jx, agent->y + jy, agent->z + jz, 0, 0);
	}
	return 0;
}

#endif
