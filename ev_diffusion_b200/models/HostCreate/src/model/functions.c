/*
 * HostCreate -- scripts for the host-side agent API fixture.
 *
 * Authored for this repository with the stock model API only: an init function
 * builds the population with h_allocate_agent_Item_array / h_add_agents_Item_default,
 * a step function adds three agents per iteration with h_add_agent_Item_default and
 * reads the population back through the analytics calls, and one agent function ages
 * the items.  No 0.xml is needed.
 */
#ifndef _FUNCTIONS_H_
#define _FUNCTIONS_H_

#include "header.h"

static int next_id = 0;

__FLAME_GPU_INIT_FUNC__ void seed_population()
{
	const unsigned int count = 50;
	xmachine_memory_Item** items = h_allocate_agent_Item_array(count);
	for (unsigned int i = 0; i < count; i++)
	{
		items[i]->id = next_id++;
		items[i]->x = 0.5f * (float)i;
		/* y and age keep their XML default values (2.5, 7) */
	}
	h_add_agents_Item_default(items, count);
	h_free_agent_Item_array(&items, count);
}

__FLAME_GPU_STEP_FUNC__ void spawn()
{
	/* the oldest item's age and the number of items decide where the newcomers start */
	const int oldest = max_Item_default_age_variable();
	const int population = get_agent_Item_default_count();
	for (int k = 0; k < 3; k++)
	{
		if (get_agent_Item_default_count() >= xmachine_memory_Item_MAX)
			break;
		xmachine_memory_Item* item = h_allocate_agent_Item();
		item->id = next_id++;
		item->x = (float)(oldest + k);
		item->crowd = population;
		h_add_agent_Item_default(item);
		h_free_agent_Item(&item);
	}
}

__FLAME_GPU_FUNC__ int grow(xmachine_memory_Item* agent)
{
	agent->age++;
	agent->x += DRIFT;
	return 0;
}

#endif
